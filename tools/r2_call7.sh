#!/bin/bash
# round 2, GPU call 7 (1 GPU): fused PNA, balanced statistics pass
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_model.py -q -x -k "pna or golden or oracle" ) > gpurun_out/r2_c7_tests.log 2>&1
tail -4 gpurun_out/r2_c7_tests.log | cut -c1-300
GNNB_FUSED_TIMING=1 timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline --steps 3 > gpurun_out/r2_c7_timing.json 2> gpurun_out/r2_c7_timing.err
grep "fused-tc phases" gpurun_out/r2_c7_timing.err | tail -1
for cp in 8 32; do
GNNB_TC_WEIGHT_COPIES=$cp timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline > gpurun_out/r2_c7_c4_copies$cp.json 2> gpurun_out/r2_c7_c4_copies$cp.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c7_c4_copies$cp.json").read().splitlines() if l.startswith("{")][-1])
print("weight copies $cp: C4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tc_kernel -s 1 -c 1 \
   -o gpurun_out/r2_fused_tc_pna_v2 python bench.py --workload c4_pna_lipo --graphs 20000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_c7_ncu.log 2>&1
tail -1 gpurun_out/r2_c7_ncu.log
