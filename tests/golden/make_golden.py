#!/usr/bin/env python
"""Generate the committed golden fixtures.  Runs only where /root/reference exists.

1. ``lib_tb/``  : the reference's own known-answer vectors for the hot path, copied verbatim
   from /root/reference/gnnbuilder/gnn_builder_lib_test/tb_data/*.bin (binary data, produced by
   the reference's gen_test_data.py with PyG; the .txt duplicates are skipped).
2. ``models/<name>.npz`` : whole-model outputs of the REFERENCE ITSELF -- ``<name>_top``
   rendered by the reference's own Project and compiled by oracle/build_ref.py -- on seeded
   synthetic graphs, for the four molecular BASELINE configs and a reduced-width variant of
   each.  Weights come from ``gnn_builder_b200.models.build_model(seed=0)``; the small variants
   store them, the full-size ones store a checksum (they are re-derived from the seed).
3. ``ref_layers.npz`` : per-layer outputs of the reference templates on a random graph with
   zero-in-degree nodes, ragged degrees and a self loop (cases tb_data does not cover).

    python tests/golden/make_golden.py
"""
import shutil
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

from gnn_builder_b200.configs import C1, C2, C3, C4  # noqa: E402
from gnn_builder_b200.data import make_molecular_batch  # noqa: E402
from gnn_builder_b200.models import build_model  # noqa: E402

import build_ref  # noqa: E402
from oracle import RefLayers, RefModel  # noqa: E402

REF_TB = Path("/root/reference/gnnbuilder/gnn_builder_lib_test/tb_data")


def params_checksum(params: dict) -> float:
    return float(sum(float(np.abs(v.astype(np.float64)).sum()) for v in params.values()))


def copy_lib_tb():
    dst = HERE / "lib_tb"
    dst.mkdir(exist_ok=True)
    n = 0
    for fp in sorted(REF_TB.glob("*.bin")):
        shutil.copyfile(fp, dst / fp.name)
        n += 1
    print(f"lib_tb: {n} files")


def model_goldens():
    out_dir = HERE / "models"
    out_dir.mkdir(exist_ok=True)
    for w in (C1, C2, C3, C4):
        for wl, n_graphs, store_params in ((build_ref.small_variant(w), 24, True), (w, 8, False)):
            build_ref.build_model(wl, pna_delta=wl.pna_delta)
            model = build_model(wl, pna_delta=wl.pna_delta, seed=0)
            params = model.named_parameter_arrays()
            ref = RefModel(wl.name)
            assert ref.param_names == list(params.keys()), (ref.param_names, list(params.keys()))
            ref.set_params(params)
            batch = make_molecular_batch(n_graphs, wl.mu_nodes, wl.mu_edges, wl.in_dim,
                                         seed=100 + wl.seed, max_nodes=min(wl.max_nodes, 60))
            out = ref.run_batch(batch)
            assert np.isfinite(out).all()
            blob = dict(x=batch.x, coo=batch.coo, node_ptr=batch.node_ptr, edge_ptr=batch.edge_ptr,
                        out=out, checksum=np.float64(params_checksum(params)))
            if store_params:
                for k, v in params.items():
                    blob["param__" + k] = v
            np.savez_compressed(out_dir / f"{wl.name}.npz", **blob)
            print(wl.name, out.shape, float(np.abs(out).max()))


def layer_goldens():
    rl = RefLayers()
    rng = np.random.default_rng(7)
    n, e, fi, fo = 40, 150, 8, 8
    src = rng.integers(0, n, e)
    dst = rng.integers(0, n - 6, e)      # nodes n-6.. have in-degree 0
    dst[:12] = 3                          # one heavy row
    src[20], dst[20] = 5, 5               # a self loop
    coo = np.stack([src, dst], 1).astype(np.int32)
    x = rng.uniform(-1, 1, (n, fi)).astype(np.float32)
    tabs = rl.tables(coo, n, with_edge_index=True)
    t4 = tabs[:4]
    W = lambda *s: rng.uniform(-0.5, 0.5, s).astype(np.float32)  # noqa: E731
    blob = dict(coo=coo, x=x, in_deg=tabs[0], out_deg=tabs[1], offsets=tabs[2], nbr=tabs[3],
                eidx=tabs[4])
    gw = [W(fo, fi), W(fo)]
    blob.update(gcn_W=gw[0], gcn_b=gw[1], gcn_out=rl.conv("gcn", x, coo, t4, gw, fo=fo))
    iw = [W(fo, fi), W(fo), W(fo, fo), W(fo)]
    blob.update(gin_W0=iw[0], gin_b0=iw[1], gin_W1=iw[2], gin_b1=iw[3],
                gin_out=rl.conv("gin", x, coo, t4, iw, scalar=0.3, fo=fo))
    sw = [W(fo, fi), W(fo), W(fo, fi)]
    blob.update(sage_Wl=sw[0], sage_bl=sw[1], sage_Wr=sw[2],
                sage_out=rl.conv("sage", x, coo, t4, sw, fo=fo))
    pw = [W(fi, 2 * fi), W(fi), W(fo, 13 * fi), W(fo), W(fo, fo), W(fo)]
    blob.update(pna_Wpre=pw[0], pna_bpre=pw[1], pna_Wpost=pw[2], pna_bpost=pw[3], pna_Wlin=pw[4],
                pna_blin=pw[5], pna_out=rl.conv("pna", x, coo, t4, pw, scalar=1.3, fo=fo))
    blob.update(lg_out=rl.same_conv("lg", x, coo, t4), simple_out=rl.same_conv("simple", x, coo, t4))
    for k in ("add", "mean", "max"):
        blob[f"pool_{k}"] = rl.pool(k, x)
    acts_in = np.linspace(-10, 10, 101).astype(np.float32)
    blob["act_in"] = acts_in
    for a in range(13):
        blob[f"act_{a}"] = rl.activation(a, acts_in)
    np.savez_compressed(HERE / "ref_layers.npz", **blob)
    print("ref_layers.npz written; pna NaN rows:", int(np.isnan(blob["pna_out"]).any(1).sum()))


if __name__ == "__main__":
    assert build_ref.have_reference(), "needs /root/reference"
    build_ref.build_all()
    copy_lib_tb()
    layer_goldens()
    model_goldens()
