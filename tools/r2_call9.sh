#!/bin/bash
# round 2, GPU call 9 (8 GPUs): the halo exchange at 4 and 8 ranks, full bench at N = 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_c9_smi.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -q ) > gpurun_out/r2_c9_tests.log 2>&1
tail -5 gpurun_out/r2_c9_tests.log | cut -c1-400
for tr in p2p nccl; do
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 8 --workload c5_gcn_large --transport $tr --steps 10 ) > gpurun_out/r2_c9_c5_$tr.json 2> gpurun_out/r2_c9_c5_$tr.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c9_c5_$tr.json").read().splitlines() if l.startswith("{")][-1])
    print("$tr N=8:", d["value"], d["ms_per_step"], "parity", d["parity"], "\n   exchange", d["exchange"])
except Exception as e:
    print("$tr failed", e); print(open("gpurun_out/r2_c9_c5_$tr.err").read()[-2000:])
PY
done
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 \
   bench.py --gpus 4 --workload c5_gcn_large --steps 10 ) > gpurun_out/r2_c9_c5_n4.json 2> gpurun_out/r2_c9_c5_n4.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c9_c5_n4.json").read().splitlines() if l.startswith("{")][-1])
    print("N=4:", d["value"], d["ms_per_step"], "parity", d["parity"]["max_rel_err"], "\n   exchange", d["exchange"])
except Exception as e:
    print("n4 failed", e)
PY
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 \
   bench.py --gpus 8 ) > gpurun_out/r2_c9_bench_n8.json 2> gpurun_out/r2_c9_bench_n8.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c9_bench_n8.json").read().splitlines() if l.startswith("{")][-1])
    print("bench N=8: C2", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("h2d_gbs_per_gpu"))
    for k,v in d.get("extra",{}).items(): print("  ", k, v.get("value"), v.get("ms_per_step"), v.get("error"), (v.get("parity") or {}).get("max_rel_err"), (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("bench n8 failed", e); print(open("gpurun_out/r2_c9_bench_n8.err").read()[-2000:])
PY
tail -3 gpurun_out/r2_c9_bench_n8.err | cut -c1-300
