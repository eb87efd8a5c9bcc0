#!/bin/bash
# round 2, GPU call 5 (1 GPU): where does the fused PNA kernel spend its time?
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -q -x -k "pna" ) > gpurun_out/r2_c5_tests.log 2>&1
tail -5 gpurun_out/r2_c5_tests.log | cut -c1-300
GNNB_FUSED_TIMING=1 timeout 300 python bench.py --workload c4_pna_lipo --no-cpu-baseline --steps 3 > gpurun_out/r2_c5_timing.json 2> gpurun_out/r2_c5_timing.err
grep "fused-tc phases" gpurun_out/r2_c5_timing.err | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tc_kernel -s 1 -c 1 \
   -o gpurun_out/r2_fused_tc_pna_v1 python bench.py --workload c4_pna_lipo --graphs 20000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_c5_ncu.log 2>&1
tail -2 gpurun_out/r2_c5_ncu.log
