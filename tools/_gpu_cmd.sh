timeout 300 python tools/sweep.py 2>&1 | grep -v "all ops"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; tail -3 gpurun_out/bench_ab.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print({k:d[k] for k in ('value','ms_per_step','overlapped','e2e','finetune')})
print(d['roofline'])
PY
