"""Live check of the C oracle against the REFERENCE ITSELF (oracle/_ref, compiled from
/root/reference by oracle/build_ref.py) on seeded random inputs, bit for bit.  Skipped where the
reference build is absent (the committed fixtures of test_oracle_golden.py still pin it)."""
import numpy as np
import pytest

from conftest import MODEL_NAMES, model_and_params
from oracle import RefLayers, RefModel, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")

DIMS = [(8, 8), (5, 12), (12, 12), (9, 64), (64, 64), (11, 128), (128, 128), (9, 128), (9, 80),
        (80, 80)]


@pytest.fixture(scope="module")
def rl():
    return RefLayers()


def _graph(rng, n, e, zero_deg=True):
    src = rng.integers(0, n, e)
    dst = rng.integers(0, max(1, n - (3 if zero_deg else 0)), e)
    return np.stack([src, dst], 1).astype(np.int32)


def test_reference_unit_testbench_passes():
    import build_ref

    if not build_ref.have_reference():
        pytest.skip("/root/reference absent")
    out = build_ref.run_lib_test()
    assert out.count("PASS") == 21 and "FAIL" not in out


@pytest.mark.parametrize("n,e", [(1, 0), (2, 1), (17, 40), (300, 1500), (1000, 900)])
def test_tables(orc, rl, n, e):
    rng = np.random.default_rng(n * 31 + e)
    coo = _graph(rng, n, e)
    ref = rl.tables(coo, n, with_edge_index=True)
    ind, outd = orc.degree_tables(coo, n)
    off, nbr, eidx = orc.neighbor_tables(coo, ind, with_edge_index=True)
    for a, b in zip((ind, outd, off, nbr, eidx), ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("fi,fo", DIMS)
@pytest.mark.parametrize("kind", ["gcn", "gin", "sage", "pna"])
def test_convs(orc, rl, kind, fi, fo):
    rng = np.random.default_rng(fi * 1000 + fo)
    n, e = 50, 170
    coo = _graph(rng, n, e, zero_deg=(kind != "pna"))
    x = rng.uniform(-1, 1, (n, fi)).astype(np.float32)
    t = rl.tables(coo, n)
    ind, outd, off, nbr = t
    W = lambda *s: rng.uniform(-0.3, 0.3, s).astype(np.float32)  # noqa: E731
    if kind == "gcn":
        w = [W(fo, fi), W(fo)]
        mine = orc.gcn_conv(x, off, nbr, ind, *w)
        ref = rl.conv(kind, x, coo, t, w, fo=fo)
    elif kind == "gin":
        w = [W(fo, fi), W(fo), W(fo, fo), W(fo)]
        mine = orc.gin_conv(x, off, nbr, ind, *w, 0.25)
        ref = rl.conv(kind, x, coo, t, w, scalar=0.25, fo=fo)
    elif kind == "sage":
        w = [W(fo, fi), W(fo), W(fo, fi)]
        mine = orc.sage_conv(x, off, nbr, ind, *w)
        ref = rl.conv(kind, x, coo, t, w, fo=fo)
    else:
        w = [W(fi, 2 * fi), W(fi), W(fo, 13 * fi), W(fo), W(fo, fo), W(fo)]
        mine = orc.pna_conv(x, off, nbr, ind, *w, 1.7)
        ref = rl.conv(kind, x, coo, t, w, scalar=1.7, fo=fo)
    assert np.array_equal(mine, ref, equal_nan=True)


@pytest.mark.parametrize("f", [8, 12, 64, 80, 128])
def test_pools(orc, rl, f):
    rng = np.random.default_rng(f)
    for n in (1, 2, 37):
        x = rng.uniform(-2, 2, (n, f)).astype(np.float32)
        for k in ("add", "mean", "max"):
            assert np.array_equal(orc.pool(k, x), rl.pool(k, x))


@pytest.mark.parametrize("fi,fo", [(10, 20), (8, 8), (384, 64), (64, 64), (64, 19), (1040, 80)])
def test_linear(orc, rl, fi, fo):
    rng = np.random.default_rng(fi + fo)
    x, W, b = (rng.uniform(-1, 1, s).astype(np.float32) for s in ((fi,), (fo, fi), (fo,)))
    y = orc.linear(x, W, b)
    assert np.array_equal(y, rl.linear(x, W, b))
    assert np.array_equal(y, rl.linear(x, W, b, buffered=True))


@pytest.mark.parametrize("name", [m + "_small" for m in MODEL_NAMES] + ["c2_gin_qm9"])
def test_whole_model_random_graphs(orc, name):
    from gnn_builder_b200.data import make_molecular_batch

    w, model, params = model_and_params(name)
    ref = RefModel(name)
    ref.set_params(params)
    batch = make_molecular_batch(12, w.mu_nodes, w.mu_edges, w.in_dim, seed=999, max_nodes=60)
    assert np.array_equal(orc.model_forward_batch(model.describe(), list(params.values()), batch),
                          ref.run_batch(batch))
