#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 0 --no-cpu-baseline > gpurun_out/r2_final2_w0.json 2> gpurun_out/r2_final2_w0.err
( time timeout 900 python bench.py ) > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err
python - <<PY
import json
for f in ("gpurun_out/r2_final2_w0.json","gpurun_out/r2_final2_bench.json"):
    d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    print(f, "C2", d["value"], "launches", d["gpu_launches"], "e2e", d["e2e"]["value"])
    for k,v in d["extra"].items(): print("   ", k, v.get("value"), v.get("gpu_launches"), v.get("error"), (v.get("parity") or {}).get("max_rel_err"))
PY
