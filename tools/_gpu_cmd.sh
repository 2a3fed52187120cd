timeout 600 python -m pytest tests/test_api_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 200 python tools/ab.py streams
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; tail -3 gpurun_out/bench_ab.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e')})
PY
