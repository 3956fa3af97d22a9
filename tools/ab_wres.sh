#!/bin/bash
# A/B on ONE box: gate-backward / dX GEMMs with resident weights (WNB_NT_WRES=1, default) vs the streaming ring (=0).
B="python bench.py --steps 20 --warmup 3 --with-decode 0 --cpu-baseline 0 --fp32-line 0 --loader-e2e 0"
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_oracle_direct.py -m gpu -q -k "not decode" 2>&1 | tail -4
for v in 0 1 0 1; do WNB_NT_WRES=$v $B > gpurun_out/ab_wres$v.json 2>gpurun_out/ab_wres$v.err; python - <<PY
import json
d=json.loads(open("gpurun_out/ab_wres$v.json").read().strip().splitlines()[-1])
print("wres=$v", round(d["ms_per_step"],3), {k["kernel"]: round(k["mean_launch_ms"]*1e3,1) for k in d.get("roofline_kernels",[])})
PY
done
