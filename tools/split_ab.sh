#!/bin/bash
# SM-budget experiment: early segment (throughput-bound layers) sized for E SMs, late segment (latency-bound) for L SMs.
mkdir -p gpurun_out
run() {
  env "$@" timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value %.4f overlapped %.4f e2e %.4f' % (d['ms_per_step'], d['overlapped']['ms_per_step'], d['e2e']['ms_per_step']), d['e2e']['regions_ms_per_step'])"
}
run KWS_NOP=1
for L in 48 64 80 96 112; do run KWS_SPLIT_AFTER=block3b_out KWS_LATE_SMS=$L; done
run KWS_SPLIT_AFTER=block3b_out KWS_EARLY_SMS=132 KWS_LATE_SMS=64
run KWS_SPLIT_AFTER=block3b_out KWS_EARLY_SMS=132 KWS_LATE_SMS=80
run KWS_LATE_SMS=64
run KWS_LATE_SMS=80
