#!/bin/bash
# round 2, GPU call 12 (2 GPUs): full GPU test suite, pipelined block sends with short-lived
# aggregation CTAs (blocks 1 / 4), C5 on one GPU with the CTA-parallel heavy-row combine
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c12_tests.log 2>&1
tail -5 gpurun_out/r2_c12_tests.log | cut -c1-400
for blk in 1 4; do
( GNNB_HALO_BLOCKS=$blk timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$blk \
   bench.py --gpus 2 --workload c5_gcn_large --transport p2p --steps 10 --no-cpu-baseline ) > gpurun_out/r2_c12_c5_blk$blk.json 2> gpurun_out/r2_c12_c5_blk$blk.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c12_c5_blk$blk.json").read().splitlines() if l.startswith("{")][-1])
    x=d["exchange"]
    print("N=2 blocks $blk:", round(d["value"]/1e9,2), "G edges/s", round(d["ms_per_step"],3), "ms; parity", d["parity"]["max_rel_err"], "| xchg alone", round(x["exchange_ms_per_layer_alone"],3), "compute alone", round(x["compute_ms_per_layer_alone"],3))
except Exception as e:
    print("blocks $blk failed", e); print(open("gpurun_out/r2_c12_c5_blk$blk.err").read()[-1500:])
PY
done
timeout 300 python bench.py --workload c5_gcn_large --no-cpu-baseline --steps 5 > gpurun_out/r2_c12_c5_n1.json 2> gpurun_out/r2_c12_c5_n1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c12_c5_n1.json").read().splitlines() if l.startswith("{")][-1])
print("N=1:", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["class_ms_per_step"], d["parity"]["max_rel_err"])
PY
