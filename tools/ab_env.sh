#!/bin/bash
# A/B on ONE box: tools/ab_env.sh VAR v1 v2 ...  -> ms/step and per-kernel launch times of the default bench workload per value
VAR=$1; shift
B="python bench.py --steps 20 --warmup 3 --with-decode 0 --cpu-baseline 0 --fp32-line 0 --loader-e2e 0"
for v in "$@" "$@"; do env $VAR=$v $B > gpurun_out/ab_$VAR$v.json 2>gpurun_out/ab_$VAR$v.err; python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$VAR$v.json").read().strip().splitlines()[-1])
print("$VAR=$v", round(d["ms_per_step"],3), {k["kernel"]: round(k["mean_launch_ms"]*1e3,1) for k in d.get("roofline_kernels",[])})
PY
done
