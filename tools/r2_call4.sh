#!/bin/bash
# round 2, GPU call 4 (1 GPU): PNA in the fused kernel -- tests, error table, bench (C4 + all extras)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_model.py -q -x -k "pna or golden or oracle or identity" ) > gpurun_out/r2_c4_tests.log 2>&1
tail -25 gpurun_out/r2_c4_tests.log | cut -c1-400
timeout 300 python tools/fused_error.py > gpurun_out/r2_c4_fused_error.txt 2>&1
cat gpurun_out/r2_c4_fused_error.txt | tail -22
( time timeout 600 python bench.py --workload c4_pna_lipo ) > gpurun_out/r2_c4_bench_c4.json 2> gpurun_out/r2_c4_bench_c4.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c4_bench_c4.json").read().splitlines() if l.startswith("{")][-1])
    print("C4 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "path", d["config"]["path"])
except Exception as e:
    print("c4 bench failed", e); print(open("gpurun_out/r2_c4_bench_c4.err").read()[-1500:])
PY
