"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gnn_builder_b200 as gnnb  # noqa: E402
from gnn_builder_b200.configs import C1, C2, C3, C4  # noqa: E402

for w in (C2, C1, C3, C4):
    model = gnnb.build_model(w, pna_delta=w.pna_delta, seed=0)
    batch = gnnb.make_molecular_batch(700, w.mu_nodes, w.mu_edges, w.in_dim, seed=1)
    with gnnb.Engine(model) as eng:
        a = eng.run(batch)
        k = eng.last_kernel
        eng.set_path(gnnb.PATH_LAYERWISE)
        b = eng.run(batch)
        import hashlib

        print(w.name, k, float(np.abs(a - b).max()), "md5(out)", hashlib.md5(a.tobytes()).hexdigest()[:12],
              flush=True)
