#!/bin/bash
# round 2, GPU call 2 (2 GPUs): full GPU test suite incl. the multi-GPU halo tests, bench at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_c2_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2_c2_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c2_tests.log 2>&1
tail -15 gpurun_out/r2_c2_tests.log
for tr in p2p nccl; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --workload c5_gcn_large --transport $tr --steps 10 ) > gpurun_out/r2_c2_c5_$tr.json 2> gpurun_out/r2_c2_c5_$tr.err
tail -c 1500 gpurun_out/r2_c2_c5_$tr.json; tail -c 800 gpurun_out/r2_c2_c5_$tr.err
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
   bench.py --gpus 2 ) > gpurun_out/r2_c2_bench_n2.json 2> gpurun_out/r2_c2_bench_n2.err
tail -c 600 gpurun_out/r2_c2_bench_n2.err
