#!/bin/bash
# A/B of an environment knob on the bench line: tools/ab_env.sh VAR  (runs bench.py with VAR unset and VAR=1)
mkdir -p gpurun_out
for V in "" 1; do
  env $1=$V timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$1_$V.json 2> gpurun_out/ab_$1_$V.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$1_$V.json"))
print("$1=$V", "value %.0f" % d["value"], "ms %.4f" % d["ms_per_step"], "overlapped %.4f" % d["overlapped"]["ms_per_step"],
      "e2e %.0f (%.4f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "finetune %.4f ms" % d["finetune"]["ms_per_step"], "launches", d["gpu_launches_per_step"], d["roofline"]["per_kernel_ms"])
PY
done
