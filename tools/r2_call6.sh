#!/bin/bash
# round 2, GPU call 6 (1 GPU): fused PNA with register neighbor lists and the 8-slot weight ring
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_layers.py -q -x ) > gpurun_out/r2_c6_tests.log 2>&1
tail -5 gpurun_out/r2_c6_tests.log | cut -c1-300
GNNB_FUSED_TIMING=1 timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline --steps 3 > gpurun_out/r2_c6_timing.json 2> gpurun_out/r2_c6_timing.err
grep "fused-tc phases" gpurun_out/r2_c6_timing.err | tail -1
for sl in 2 3; do
GNNB_TC_RING_SLOTS_LOG2=$sl timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline > gpurun_out/r2_c6_c4_slots$sl.json 2> gpurun_out/r2_c6_c4_slots$sl.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c6_c4_slots$sl.json").read().splitlines() if l.startswith("{")][-1])
print("slots log2 $sl: C4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_c6_c2.json 2> gpurun_out/r2_c6_c2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c6_c2.json").read().splitlines() if l.startswith("{")][-1])
print("C2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
