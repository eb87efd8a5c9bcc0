#!/bin/bash
# round 2, GPU call 1 (1 GPU): tests, unexecuted probes, bench with extras, hub-L2 sweep on C5
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_c1_tests.log 2>&1
tail -5 gpurun_out/r2_c1_tests.log
timeout 120 python tools/mma_rate.py > gpurun_out/r2_c1_mma_rate.txt 2>&1
timeout 120 python tools/tmem_bf16_probe.py > gpurun_out/r2_c1_tmem_bf16.txt 2>&1
tail -9 gpurun_out/r2_c1_tmem_bf16.txt
( time timeout 900 python bench.py ) > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err
tail -c 600 gpurun_out/r2_c1_bench.err
for mb in 0 16 32 48 64 96; do
  GNNB_HUB_L2_MB=$mb timeout 300 python bench.py --workload c5_gcn_large --no-cpu-baseline --steps 5 \
    > gpurun_out/r2_c1_c5_hub$mb.json 2> gpurun_out/r2_c1_c5_hub$mb.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_c1_c5_hub$mb.json").read().strip().splitlines()[-1])
    print("hub $mb MB:", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["class_ms_per_step"], d.get("parity"))
except Exception as e:
    print("hub $mb failed", e)
PY
done
