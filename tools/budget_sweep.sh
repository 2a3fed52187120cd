#!/bin/bash
run() {
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'overlapped %.4f e2e %.4f' % (d['overlapped']['ms_per_step'], d['e2e']['ms_per_step']), d['e2e']['regions_ms_per_step'])"
}
for b in 132,80 140,72 124,88 148,80 132,64 120,96 140,96 116,72; do run --sm-budget $b; done
run --sm-budget 132,80 --pipe-streams 2
run --sm-budget 132,80 --pipe-streams 4
run --sm-budget 124,88 --pipe-streams 4
