#!/bin/bash
# round 2, GPU call 14 (1 GPU): fused PNA with the atom-major group issue
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py -q -x ) > gpurun_out/r2_c14_tests.log 2>&1
tail -4 gpurun_out/r2_c14_tests.log | cut -c1-300
GNNB_FUSED_TIMING=1 timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline --steps 3 > gpurun_out/r2_c14_timing.json 2> gpurun_out/r2_c14_timing.err
grep "fused-tc" gpurun_out/r2_c14_timing.err | tail -2
timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline > gpurun_out/r2_c14_c4.json 2> gpurun_out/r2_c14_c4.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c14_c4.json").read().splitlines() if l.startswith("{")][-1])
print("C4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
timeout 200 python tools/fused_error.py 2>&1 | tail -8
