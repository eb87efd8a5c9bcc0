#!/bin/bash
# Round-2 record run on an 8-GPU box (gpurun --gpus 8 -- bash tools/record_run_ngpu.sh): what the driver runs at round end -- multi-GPU tests, then the
# default bench at N = 8, 4, 2 (C2 + extra.c4 + extra.c5 in one line each)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_model.py -q -k "multi or halo or beyond" ) > gpurun_out/r2_c16_tests.log 2>&1
tail -4 gpurun_out/r2_c16_tests.log | cut -c1-300
for n in 8 4 2; do
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n \
   bench.py --gpus $n ) > gpurun_out/r2_c16_bench_n$n.json 2> gpurun_out/r2_c16_bench_n$n.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_c16_bench_n$n.json").read().splitlines() if l.startswith("{")][-1])
    print("N=$n: C2", round(d["value"]/1e6,1), "M e2e", round(d["e2e"]["value"]/1e6,1), "M h2d/gpu", round(d["e2e"].get("h2d_gbs_per_gpu",0),1))
    for k,v in d.get("extra",{}).items():
        x=v.get("exchange") or {}
        print("   ", k, v.get("value"), v.get("ms_per_step"), "e2e", (v.get("e2e") or {}).get("value"), v.get("error"), (v.get("parity") or {}).get("max_rel_err"), x.get("transport"), x.get("exchange_ms_per_layer_alone"))
except Exception as e:
    print("bench N=$n failed", e); print(open("gpurun_out/r2_c16_bench_n$n.err").read()[-1500:])
PY
tail -4 gpurun_out/r2_c16_bench_n$n.err | grep real
done
