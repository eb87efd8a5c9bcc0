#!/bin/bash
# round 2, GPU call 8 (1 GPU): fused PNA v3 (output pass writes the next A operand) + issuer timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py -q -x ) > gpurun_out/r2_c8_tests.log 2>&1
tail -4 gpurun_out/r2_c8_tests.log | cut -c1-300
GNNB_FUSED_TIMING=1 timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline --steps 3 > gpurun_out/r2_c8_timing.json 2> gpurun_out/r2_c8_timing.err
grep "fused-tc" gpurun_out/r2_c8_timing.err | tail -2
GNNB_FUSED_TIMING=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 3 > gpurun_out/r2_c8_timing_c2.json 2> gpurun_out/r2_c8_timing_c2.err
grep "fused-tc" gpurun_out/r2_c8_timing_c2.err | tail -2
timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline > gpurun_out/r2_c8_c4.json 2> gpurun_out/r2_c8_c4.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c8_c4.json").read().splitlines() if l.startswith("{")][-1])
print("C4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
