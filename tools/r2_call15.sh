#!/bin/bash
# round 2, GPU call 15 (1 GPU): full GPU suite + the default bench line on the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c15_tests.log 2>&1
tail -4 gpurun_out/r2_c15_tests.log | cut -c1-300
GNNB_FUSED_TIMING=1 timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline --steps 3 > gpurun_out/r2_c15_timing.json 2> gpurun_out/r2_c15_timing.err
grep "fused-tc" gpurun_out/r2_c15_timing.err | tail -2
( time timeout 900 python bench.py ) > gpurun_out/r2_c15_bench.json 2> gpurun_out/r2_c15_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c15_bench.json").read().splitlines() if l.startswith("{")][-1])
print("C2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for k,v in d["extra"].items(): print(k, v.get("value"), v.get("ms_per_step"), "e2e", (v.get("e2e") or {}).get("value"), v.get("error"), (v.get("parity") or {}).get("max_rel_err"))
PY
tail -3 gpurun_out/r2_c15_bench.err | cut -c1-200
