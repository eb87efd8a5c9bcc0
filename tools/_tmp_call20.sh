#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c20_tests.log 2>&1
tail -2 gpurun_out/r2_c20_tests.log | cut -c1-300
for cfg in "c1_gcn_esol 1000000" "c3_sage_hiv 500000"; do
set -- $cfg
timeout 300 python bench.py --workload $1 --graphs $2 --no-extras --cpu-baseline-graphs 2000 > gpurun_out/r2_c20_$1.json 2> gpurun_out/r2_c20_$1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_c20_$1.json").read().splitlines() if l.startswith("{")][-1])
print("$1", round(d["value"]/1e6,2), "M graphs/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"]/1e6,2), "M; cpu 1 core", round(d["cpu_baseline"]["value"]), "; path", d["config"]["path"])
PY
done
