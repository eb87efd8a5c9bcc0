"""GPU parity, layer level: every C-ABI layer function against the CPU oracle (bit-exact with
the reference, see test_oracle_*.py) and against the reference's golden vectors.

Tolerances:  integer tables ==;  fp32 FAST mode max|d| <= 1e-4 * max(1, max|ref|);
STRICT mode (reference operation order, no FMA contraction) == for everything built from
+,*,/,sqrt,max (GCN/GIN/SAGE, linear, pools)."""
import numpy as np
import pytest

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def L():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from gnn_builder_b200 import layers

    return layers


def rand_graph(rng, n, e, zero_deg=True):
    src = rng.integers(0, n, e)
    dst = rng.integers(0, max(1, n - (3 if zero_deg and n > 3 else 0)), e)
    return np.stack([src, dst], 1).astype(np.int32)


# ------------------------------------------------------------------------------- tables
def test_tables_reference_fixture(L, lib_tb):
    ind, outd = L.compute_degree_tables(lib_tb.coo, lib_tb.n)
    assert np.array_equal(ind, lib_tb.in_deg) and np.array_equal(outd, lib_tb.out_deg)
    off, nbr, eidx = L.compute_neighbor_tables(lib_tb.coo, ind, outd, with_edge_index=True)
    assert np.array_equal(off, lib_tb.offsets)
    assert np.array_equal(nbr, lib_tb.nbr)
    assert np.array_equal(eidx, lib_tb.eidx)
    off2, nbr2 = L.compute_neighbor_tables(lib_tb.coo, ind, outd)
    assert np.array_equal(off2, off) and np.array_equal(nbr2, nbr)


@pytest.mark.parametrize("n,e", [(1, 0), (1, 3), (2, 1), (17, 40), (1000, 900), (5000, 200000),
                                 (300000, 2000000)])
def test_tables_random_bit_exact(L, orc, n, e):
    rng = np.random.default_rng(n + e)
    coo = rand_graph(rng, n, e)
    if e > 1000:  # hubs: many duplicates of one destination stress the stable ordering
        coo[: e // 10, 1] = 7 % n
    ind, outd = L.compute_degree_tables(coo, n)
    rind, routd = orc.degree_tables(coo, n)
    assert np.array_equal(ind, rind) and np.array_equal(outd, routd)
    off, nbr, eidx = L.compute_neighbor_tables(coo, ind, outd, with_edge_index=True)
    roff, rnbr, reidx = orc.neighbor_tables(coo, rind, with_edge_index=True)
    assert np.array_equal(off, roff) and np.array_equal(nbr, rnbr) and np.array_equal(eidx, reidx)


def test_tables_device_pointers(L, orc):
    import torch

    rng = np.random.default_rng(5)
    coo = rand_graph(rng, 500, 3000)
    dcoo = torch.from_numpy(coo).cuda()
    ind, outd = L.compute_degree_tables(dcoo, 500)
    off, nbr = L.compute_neighbor_tables(dcoo, ind, outd)
    r = orc.tables(coo, 500)
    for a, b in zip((ind, outd, off, nbr), r):
        assert np.array_equal(a.cpu().numpy(), b)


# ------------------------------------------------------------------------------- linear / act
@pytest.mark.parametrize("rows,fi,fo", [(1, 10, 20), (7, 9, 64), (300, 128, 128), (129, 384, 64),
                                        (64, 64, 19), (33, 1040, 80), (5, 3, 1),
                                        # >= 512 rows: the tcgen05 (3xTF32) GEMM of gemm_tc.cu
                                        (512, 128, 128), (5000, 128, 128), (4099, 1040, 80),
                                        (3001, 11, 160), (7777, 80, 19), (1000, 33, 8),
                                        (20000, 384, 64), (600, 200, 192)])
def test_linear(L, orc, rows, fi, fo):
    rng = np.random.default_rng(rows * fi + fo)
    x = rng.uniform(-1, 1, (rows, fi)).astype(np.float32)
    W = rng.uniform(-1, 1, (fo, fi)).astype(np.float32)
    b = rng.uniform(-1, 1, fo).astype(np.float32)
    ref = np.stack([orc.linear(x[i], W, b) for i in range(rows)])
    assert rel_err(L.linear(x, W, b), ref) < TOL
    assert np.array_equal(L.linear(x, W, b, math=L.STRICT), ref)


def test_linear_closed_form_exact(L):
    # the reference's own linear test uses integer-valued data and `!=` (test.cpp:678-745)
    x = np.arange(10, dtype=np.float32)
    W = (np.arange(20)[:, None] + np.arange(10)[None, :]).astype(np.float32)
    b = np.arange(20, dtype=np.float32)
    assert np.array_equal(L.linear(x, W, b), W @ x + b)


ACT_NAMES = {0: "identity", 1: "relu", 2: "gelu_approx_tanh", 3: "sigmoid", 4: "tanh", 5: "elu",
             6: "hardtanh", 7: "leakyrelu", 8: "gelu", 9: "silu", 10: "softsign", 11: "sin",
             12: "cos"}


@pytest.mark.parametrize("act", sorted(ACT_NAMES))
def test_activations(L, orc, lib_tb, act):
    x = lib_tb.f32(f"test_activations_x_in_{ACT_NAMES[act]}")
    gold = lib_tb.f32(f"test_activations_x_out_{ACT_NAMES[act]}")
    y = L.apply_activation(act, x)
    assert np.abs(y - gold).max() < 1e-3            # the reference's own tolerance (test.cpp:14)
    assert np.abs(y - orc.activation(act, x)).max() < 2e-6
    z = np.load(GOLDEN / "ref_layers.npz")
    assert np.abs(L.apply_activation(act, z["act_in"]) - z[f"act_{act}"]).max() < 2e-6


# ------------------------------------------------------------------------------- convs
def _conv_all(L, orc, x, coo, n, seed, fo, strict_too=True, pna=True):
    rng = np.random.default_rng(seed)
    fi = x.shape[1]
    ind, outd, off, nbr = orc.tables(coo, n)
    W = lambda *s: rng.uniform(-0.3, 0.3, s).astype(np.float32)  # noqa: E731
    res = {}
    w = [W(fo, fi), W(fo)]
    ref = orc.gcn_conv(x, off, nbr, ind, *w)
    res["gcn"] = (L.gcn_conv(x, coo, off, nbr, ind, outd, *w), ref)
    if strict_too:
        assert np.array_equal(L.gcn_conv(x, coo, off, nbr, ind, outd, *w, math=L.STRICT), ref)
    w = [W(fo, fi), W(fo), W(fo, fo), W(fo)]
    ref = orc.gin_conv(x, off, nbr, ind, *w, 0.25)
    res["gin"] = (L.gin_conv(x, coo, off, nbr, ind, outd, *w, 0.25), ref)
    if strict_too:
        assert np.array_equal(L.gin_conv(x, coo, off, nbr, ind, outd, *w, 0.25, math=L.STRICT), ref)
    w = [W(fo, fi), W(fo), W(fo, fi)]
    ref = orc.sage_conv(x, off, nbr, ind, *w)
    res["sage"] = (L.sage_conv(x, coo, off, nbr, ind, outd, *w), ref)
    if strict_too:
        assert np.array_equal(L.sage_conv(x, coo, off, nbr, ind, outd, *w, math=L.STRICT), ref)
    if pna:
        w = [W(fi, 2 * fi), W(fi), W(fo, 13 * fi), W(fo), W(fo, fo), W(fo)]
        ref = orc.pna_conv(x, off, nbr, ind, *w, 1.7)
        got = L.pna_conv(x, coo, off, nbr, ind, outd, *w, 1.7)
        assert np.array_equal(np.isnan(got), np.isnan(ref))   # NaN rows = zero in-degree rows
        ok = ~np.isnan(ref)
        res["pna"] = (got[ok], ref[ok])
    for k, (got, ref) in res.items():
        assert rel_err(got, ref) < TOL, (k, rel_err(got, ref))


def test_convs_reference_fixture_vs_pyg_golden(L, lib_tb):
    """The reference's own conv tests (test.cpp:1056-1726) with its tolerances, on the GPU."""
    t = lib_tb
    y = L.gcn_conv(t.x, t.coo, t.offsets, t.nbr, t.in_deg, t.out_deg,
                   t.f32("tb_gcn_weights", 8, 8), t.f32("tb_gcn_bias"))
    assert np.abs(y - t.f32("tb_gcn_output", t.n, 8)).max() < 1e-5
    y = L.gin_conv(t.x, t.coo, t.offsets, t.nbr, t.in_deg, t.out_deg,
                   t.f32("tb_gin_mlp_0_weights", 8, 8), t.f32("tb_gin_mlp_0_bias"),
                   t.f32("tb_gin_mlp_1_weights", 8, 8), t.f32("tb_gin_mlp_1_bias"),
                   float(t.f32("tb_gin_eps")[0]))
    assert np.abs(y - t.f32("tb_gin_output", t.n, 8)).max() < 1e-5
    y = L.sage_conv(t.x, t.coo, t.offsets, t.nbr, t.in_deg, t.out_deg,
                    t.f32("tb_sage_neighbor_lin_weights", 8, 8),
                    t.f32("tb_sage_neighbor_lin_bias"), t.f32("tb_sage_self_lin_weights", 8, 8))
    assert np.abs(y - t.f32("tb_sage_output", t.n, 8)).max() < 1e-5
    y = L.pna_conv(t.x, t.coo, t.offsets, t.nbr, t.in_deg, t.out_deg,
                   t.f32("tb_pna_transform_lin_weights", 8, 16), t.f32("tb_pna_transform_lin_bias"),
                   t.f32("tb_pna_apply_lin_weights", 8, 104), t.f32("tb_pna_apply_lin_bias"),
                   t.f32("tb_pna_final_lin_weights", 8, 8), t.f32("tb_pna_final_lin_bias"),
                   float(t.f32("tb_pna_avg_degree_log")[0]))
    assert np.abs(y - t.f32("tb_pna_output", t.n, 8)).max() < 1e-2   # test.cpp:1591


def test_convs_reference_fixture_vs_oracle(L, orc, lib_tb):
    _conv_all(L, orc, lib_tb.x, lib_tb.coo, lib_tb.n, 1, 8)


def test_convs_edge_cases_from_reference_templates(L, orc):
    """zero in-degree rows, a heavy row, a self loop: fixture produced by the compiled reference"""
    z = np.load(GOLDEN / "ref_layers.npz")
    x, coo = z["x"], z["coo"]
    n = x.shape[0]
    t = (z["offsets"], z["nbr"], z["in_deg"], z["out_deg"])
    y = L.gcn_conv(x, coo, t[0], t[1], t[2], t[3], z["gcn_W"], z["gcn_b"], math=L.STRICT)
    assert np.array_equal(y, z["gcn_out"])
    y = L.gin_conv(x, coo, t[0], t[1], t[2], t[3], z["gin_W0"], z["gin_b0"], z["gin_W1"],
                   z["gin_b1"], 0.3, math=L.STRICT)
    assert np.array_equal(y, z["gin_out"])
    y = L.sage_conv(x, coo, t[0], t[1], t[2], t[3], z["sage_Wl"], z["sage_bl"], z["sage_Wr"],
                    math=L.STRICT)
    assert np.array_equal(y, z["sage_out"])
    y = L.pna_conv(x, coo, t[0], t[1], t[2], t[3], z["pna_Wpre"], z["pna_bpre"], z["pna_Wpost"],
                   z["pna_bpost"], z["pna_Wlin"], z["pna_blin"], 1.3)
    assert np.array_equal(np.isnan(y), np.isnan(z["pna_out"]))
    ok = ~np.isnan(z["pna_out"])
    assert rel_err(y[ok], z["pna_out"][ok]) < TOL
    for k, fn in (("add", L.global_add_pool), ("mean", L.global_mean_pool), ("max", L.global_max_pool)):
        assert np.array_equal(fn(x), z[f"pool_{k}"])
    _conv_all(L, orc, x, coo, n, 2, 8)


# ------------------------------------------------------------------------------- gine / lg / simple
def test_gine_lg_simple_reference_fixture_vs_pyg_golden(L, lib_tb):
    """the reference's own tests of the three remaining convs (test.cpp:1287-1455, 1728-1919)"""
    t = lib_tb
    ef = t.f32("tb_input_edge_features", t.e, 16)
    y = L.gine_conv(t.x, ef, t.coo, t.offsets, t.nbr, t.eidx, t.in_deg, t.out_deg,
                    t.f32("tb_gine_edge_proj_weights", 8, 16), t.f32("tb_gine_edge_proj_bias"),
                    t.f32("tb_gine_mlp_0_weights", 8, 8), t.f32("tb_gine_mlp_0_bias"),
                    t.f32("tb_gine_mlp_1_weights", 8, 8), t.f32("tb_gine_mlp_1_bias"),
                    float(t.f32("tb_gine_eps")[0]))
    assert np.abs(y - t.f32("tb_gine_output", t.n, 8)).max() < 1e-5
    y = L.lg_conv(t.x, t.coo, t.offsets, t.nbr, t.in_deg, t.out_deg)
    assert np.abs(y - t.f32("tb_lgconv_output", t.n, 8)).max() < 1e-5
    y = L.simple_conv(t.x, t.coo, t.offsets, t.nbr, t.in_deg, t.out_deg)
    assert np.abs(y - t.f32("tb_simple_output", t.n, 8)).max() < 1e-5


@pytest.mark.parametrize("n,e,fi,fo,fe", [(1, 0, 8, 8, 4), (2, 3, 8, 16, 16), (100, 534, 8, 8, 16),
                                          (700, 3000, 128, 128, 32), (300, 2000, 30, 20, 7),
                                          (5000, 60000, 64, 64, 16)])
def test_gine_lg_simple_random(L, orc, n, e, fi, fo, fe):
    rng = np.random.default_rng(n + 3 * e + fi)
    coo = rand_graph(rng, n, e)
    if e > 10000:
        coo[: e // 8, 1] = 5      # a heavy destination row
    x = rng.uniform(-1, 1, (n, fi)).astype(np.float32)
    ef = rng.uniform(-1, 1, (e, fe)).astype(np.float32)
    ind, outd = orc.degree_tables(coo, n)
    off, nbr, eidx = orc.neighbor_tables(coo, ind, with_edge_index=True)
    W = lambda *s: rng.uniform(-0.3, 0.3, s).astype(np.float32)  # noqa: E731
    w = [W(fi, fe), W(fi), W(fo, fi), W(fo), W(fo, fo), W(fo)]
    ref = orc.gine_conv(x, ef, off, nbr, eidx, ind, *w, 0.2)
    got = L.gine_conv(x, ef, coo, off, nbr, eidx, ind, outd, *w, 0.2)
    assert rel_err(got, ref) < TOL
    assert np.array_equal(L.gine_conv(x, ef, coo, off, nbr, eidx, ind, outd, *w, 0.2, math=L.STRICT), ref)
    for fn, ofn in ((L.lg_conv, orc.lg_conv), (L.simple_conv, orc.simple_conv)):
        ref = ofn(x, off, nbr, ind)
        got = fn(x, coo, off, nbr, ind, outd)
        with np.errstate(invalid="ignore"):
            assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert rel_err(np.nan_to_num(got), np.nan_to_num(ref)) < TOL
        assert np.array_equal(fn(x, coo, off, nbr, ind, outd, math=L.STRICT), ref, equal_nan=True)


@pytest.mark.parametrize("n,e,fi,fo", [(1, 0, 9, 64), (2, 1, 11, 128), (40, 90, 64, 64),
                                       (700, 3000, 128, 128), (300, 900, 80, 80), (50, 200, 5, 12),
                                       (3000, 50000, 32, 16)])
def test_convs_random(L, orc, n, e, fi, fo):
    rng = np.random.default_rng(n * 7 + fi)
    coo = rand_graph(rng, n, e)
    x = rng.uniform(-1, 1, (n, fi)).astype(np.float32)
    _conv_all(L, orc, x, coo, n, 3, fo, pna=(fi <= 80))


def test_heavy_rows_and_large_graph(L, orc):
    """power-law graph: a few destination rows with thousands of neighbors"""
    from gnn_builder_b200.data import make_powerlaw_graph

    n = 20000
    x, coo = make_powerlaw_graph(n, 16, 128, seed=9, max_degree=6000)
    coo[:5000, 1] = 11      # force one row above the heavy threshold
    ind, outd, off, nbr = orc.tables(coo, n)
    assert ind.max() > 1024
    W = np.random.default_rng(0).uniform(-0.1, 0.1, (128, 128)).astype(np.float32)
    b = np.zeros(128, np.float32)
    ref = orc.gcn_conv(x, off, nbr, ind, W, b)
    got = L.gcn_conv(x, coo, off, nbr, ind, outd, W, b)
    assert rel_err(got, ref) < TOL


@pytest.mark.parametrize("n,f", [(0, 8), (1, 8), (37, 12), (600, 128), (100000, 64)])
def test_pools(L, orc, n, f):
    rng = np.random.default_rng(n + f)
    x = rng.uniform(-2, 2, (n, f)).astype(np.float32)
    for k, fn in (("add", L.global_add_pool), ("mean", L.global_mean_pool), ("max", L.global_max_pool)):
        ref = orc.pool(k, x)
        got = fn(x)
        if n <= 8192:
            assert np.array_equal(got, ref), k       # same order of additions
        else:
            assert rel_err(got, ref) < TOL, k
