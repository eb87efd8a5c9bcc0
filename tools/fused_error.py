"""Measured error of the fused tcgen05 kernel against the CPU oracle (the reference restated in
C): max |out - ref| / max(1, max |ref|) over a seeded batch, for the BASELINE configs the kernel
covers and the variants of tests/test_gpu_model.py.  The bound the tests enforce is 1e-4."""
import dataclasses
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))
import gnn_builder_b200 as gnnb  # noqa: E402
from conftest import rel_err, workload_by_name  # noqa: E402
from oracle import Oracle  # noqa: E402
from test_gpu_model import PNA_VARIANTS, VARIANTS  # noqa: E402

orc = Oracle()
cases = [(n, n, {}) for n in ("c1_gcn_esol", "c2_gin_qm9", "c3_sage_hiv", "c4_pna_lipo")]
cases += [(k, b, o) for k, (b, o) in sorted(VARIANTS.items())]
cases += [(k, "c4_pna_lipo", o) for k, o in sorted(PNA_VARIANTS.items())]
worst = 0.0
for label, base, over in cases:
    w = dataclasses.replace(workload_by_name(base), **over)
    model = gnnb.build_model(w, pna_delta=w.pna_delta, seed=11)
    params = model.named_parameter_arrays()
    batch = gnnb.make_molecular_batch(2000, w.mu_nodes, w.mu_edges, w.in_dim, seed=5)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model) as eng:
        out = eng.run(batch)
        kern = eng.last_kernel
        eng.set_path(gnnb.PATH_LAYERWISE)
        lw = eng.run(batch)
    e, e_lw = rel_err(out, ref), rel_err(lw, ref)
    worst = max(worst, e)
    print(f"{label:28s} {kern:14s} fused err {e:.2e}   layerwise (3xTF32) err {e_lw:.2e}", flush=True)
print(f"worst fused error {worst:.2e} (bound enforced by the tests: 1e-4)")
