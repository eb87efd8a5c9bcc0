#!/bin/bash
# round 2, GPU call 11 (2 GPUs): pipelined block sends -- multi-GPU tests + C5 at N = 2, blocks 1 vs 4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_large.py -q ) > gpurun_out/r2_c11_tests.log 2>&1
tail -5 gpurun_out/r2_c11_tests.log | cut -c1-400
for blk in 1 4 8; do
( GNNB_HALO_BLOCKS=$blk timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$blk \
   bench.py --gpus 2 --workload c5_gcn_large --transport p2p --steps 10 --no-cpu-baseline ) > gpurun_out/r2_c11_c5_blk$blk.json 2> gpurun_out/r2_c11_c5_blk$blk.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c11_c5_blk$blk.json").read().splitlines() if l.startswith("{")][-1])
    x=d["exchange"]
    print("N=2 blocks $blk:", round(d["value"]/1e9,2), "G edges/s", round(d["ms_per_step"],3), "ms; parity", d["parity"]["max_rel_err"], "| xchg alone", round(x["exchange_ms_per_layer_alone"],3), "compute alone", round(x["compute_ms_per_layer_alone"],3))
except Exception as e:
    print("blocks $blk failed", e); print(open("gpurun_out/r2_c11_c5_blk$blk.err").read()[-1500:])
PY
done
