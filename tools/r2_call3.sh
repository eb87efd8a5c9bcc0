#!/bin/bash
# round 2, GPU call 3 (1 GPU): bf16x2 fused kernel -- parity tests, error measurement, bench, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_tc.py -q -x ) > gpurun_out/r2_c3_tests.log 2>&1
tail -12 gpurun_out/r2_c3_tests.log
timeout 300 python tools/fused_error.py > gpurun_out/r2_c3_fused_error.txt 2>&1
cat gpurun_out/r2_c3_fused_error.txt
( time timeout 600 python bench.py --no-extras ) > gpurun_out/r2_c3_bench.json 2> gpurun_out/r2_c3_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_c3_bench.json").read().splitlines() if l.startswith("{")][-1])
print("C2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tc_kernel -s 1 -c 1 \
   -o gpurun_out/r2_fused_tc_bf2 python bench.py --graphs 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_c3_ncu.log 2>&1
tail -3 gpurun_out/r2_c3_ncu.log
