#!/bin/bash
# round 2, GPU call 10 (8 GPUs): halo exchange with peer-interleaved pack, transport autotune
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for cfg in "8 p2p" "8 nccl" "8 auto" "4 auto" "2 auto"; do
set -- $cfg
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2955$1 \
   bench.py --gpus $1 --workload c5_gcn_large --transport $2 --steps 10 --no-cpu-baseline ) > gpurun_out/r2_c10_c5_n$1_$2.json 2> gpurun_out/r2_c10_c5_n$1_$2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c10_c5_n$1_$2.json").read().splitlines() if l.startswith("{")][-1])
    x=d["exchange"]
    print("N=$1 $2:", round(d["value"]/1e9,2), "G edges/s", round(d["ms_per_step"],3), "ms; parity", d["parity"]["max_rel_err"], "| transport", x["transport"], x.get("transport_autotune_ms"), "xchg alone", round(x["exchange_ms_per_layer_alone"],3), "ms", round(x["exchange_gbs_in_per_gpu"]), "GB/s; compute alone", round(x["compute_ms_per_layer_alone"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
except Exception as e:
    print("N=$1 $2 failed", e); print(open("gpurun_out/r2_c10_c5_n$1_$2.err").read()[-1500:])
PY
done
