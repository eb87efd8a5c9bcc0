"""GPU parity for the large-graph path (BASELINE config 5 shape, scaled to what the CPU oracle
finishes in seconds): the layerwise kernels with degree-bucketed aggregation through the model
handle, and the row-partitioned runner (world size 1 here; world size 2 is exercised by
tests/run_large_multi_gpu.py under torchrun on a 2-GPU box)."""
import dataclasses

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def setup():
    import torch

    assert torch.cuda.is_available()
    import gnn_builder_b200 as gnnb
    from gnn_builder_b200.configs import C5

    w = dataclasses.replace(C5, out_dim=16)
    model = gnnb.build_model(w, seed=1)
    n = 40000
    x, coo = gnnb.make_powerlaw_graph(n, 16, w.in_dim, seed=5, max_degree=5000)
    coo[:3000, 1] = 17      # a heavy row that one CTA finishes by itself (256 < in-degree <= 4096)
    coo[3000:9500, 1] = 23  # a hub whose neighbor list is sliced over several CTAs (> 4096)
    return gnnb, w, model, n, x, coo


def test_large_graph_through_model_handle(setup, orc):
    gnnb, w, model, n, x, coo = setup
    params = list(model.named_parameter_arrays().values())
    ref, ref_emb = orc.model_forward(model.describe(), params, x, coo, return_node_emb=True)
    with gnnb.Engine(model) as eng:
        out = eng.run_graph(x, coo)
        assert eng.last_kernel == "layerwise"
        emb = eng.node_embeddings(n)
    assert rel_err(emb, ref_emb) < TOL
    assert rel_err(out, ref) < TOL


def test_row_partition_runner_world_1(setup, orc):
    import torch

    gnnb, w, model, n, x, coo = setup
    from gnn_builder_b200.distributed import LargeGraphGCN, RowPartition

    params = list(model.named_parameter_arrays().values())
    ref, ref_emb = orc.model_forward(model.describe(), params, x, coo, return_node_emb=True)
    runner = LargeGraphGCN(model, n, 0, 1).setup(RowPartition(n, 1).local_edges(coo, 0))
    out, emb = runner.forward(x, return_embeddings=True)
    torch.cuda.synchronize()
    assert rel_err(emb.cpu().numpy(), ref_emb) < TOL
    assert rel_err(out.cpu().numpy(), ref) < TOL
    # tables of the partition slice are the reference's tables, bit for bit
    ind, outd, off, nbr = orc.tables(coo, n)
    t_ind, t_off, t_nbr = runner.tables
    assert np.array_equal(t_ind.cpu().numpy(), ind)
    assert np.array_equal(t_off.cpu().numpy(), off)
    assert np.array_equal(t_nbr.cpu().numpy()[: coo.shape[0]], nbr)


# ------------------------------------------------------------------------------------------------
# the halo-exchange building blocks on ONE GPU (the multi-GPU run is tests/test_gpu_multi.py)
def _p(t):
    import ctypes as C

    return C.c_void_p(t.data_ptr()) if t is not None else None


def _gcn_rows_ref(orc, x, coo, n, W, b):
    """reference gcn_conv + ReLU for every row (oracle restatement of lib:1291-1387)"""
    ind, _, off, nbr = orc.tables(coo, n)
    return np.maximum(orc.gcn_conv(x, off, nbr, ind, W, b), 0.0)


def test_partition_layer_wider_second_call_with_hub_row(setup, orc):
    """the heavy-row scratch must follow the widest layer: emb_in 8 first, then
    64, on a partition that has rows above the heavy threshold"""
    import ctypes as C
    import torch

    gnnb, w, model, n, x, coo = setup
    from gnn_builder_b200 import _lib

    lib = _lib.load()
    n2 = 6000
    rng = np.random.default_rng(3)
    coo2 = np.stack([rng.integers(0, n2, 60000), rng.integers(0, n2, 60000)], 1).astype(np.int32)
    coo2[:5000, 1] = 11                     # in-degree 5000 > heavy threshold (256), sliced over CTAs
    dcoo = torch.from_numpy(coo2).cuda()
    ind = torch.empty(n2, dtype=torch.int32, device="cuda")
    off = torch.empty(n2, dtype=torch.int32, device="cuda")
    nbr = torch.empty(coo2.shape[0], dtype=torch.int32, device="cuda")
    _lib.check(lib.gnnb_partition_tables(_p(dcoo), 0, n2, coo2.shape[0], _p(ind), _p(off), _p(nbr), None))
    dinv = torch.empty(n2, device="cuda")
    _lib.check(lib.gnnb_degree_inv_sqrt(_p(ind), _p(dinv), n2, None))
    for fi, fo in ((8, 64), (64, 32)):
        xs = rng.uniform(-1, 1, (n2, fi)).astype(np.float32)
        Wm = rng.uniform(-0.3, 0.3, (fo, fi)).astype(np.float32)
        bm = rng.uniform(-0.1, 0.1, fo).astype(np.float32)
        y = torch.empty((n2, fo), device="cuda")
        dxs, dW, db = (torch.from_numpy(a).cuda() for a in (xs, Wm, bm))   # (kept alive over the call)
        _lib.check(lib.gnnb_gcn_conv_partition(
            n2, 0, n2, coo2.shape[0], _p(dxs), _p(y), _p(off), _p(nbr), _p(ind), _p(dinv), _p(dW),
            _p(db), None, fi, fo, 1, None))
        torch.cuda.synchronize()
        assert rel_err(y.cpu().numpy(), _gcn_rows_ref(orc, xs, coo2, n2, Wm, bm)) < TOL, (fi, fo)


@pytest.mark.parametrize("hub_mb", [0, 1])
def test_halo_plan_split_csr_matches_reference(setup, orc, hub_mb):
    """rank 1 of a 2-way row partition emulated on one GPU: ext index space, split CSR, phase 1
    (owned-source edges) + phase 2 (halo-source edges, normalise, transform) = the reference's
    gcn_conv rows; with and without hub-source L2 hints (same bits either way)"""
    import torch

    gnnb, w, model, n, x, coo = setup
    from gnn_builder_b200.distributed import CudaBackend, HaloPlan, RowPartition

    B = CudaBackend()
    part = RowPartition(n, 2)
    rank = 1
    r0, r1 = part.rows(rank)
    ind, off, nbr = B.partition_tables(B.to_device(part.local_edges(coo, rank)), r0, part.n_local)
    plan = HaloPlan(ind, nbr, rank, part, dist=None)
    refs = np.unique(part.local_edges(coo, rank)[:, 0])
    assert np.array_equal(plan.halo_ids.cpu().numpy(), refs[(refs < r0) | (refs >= r1)])
    ind_full, _, _, _ = orc.tables(coo, n)
    dinv_full = B.dinv(B.to_device(ind_full))
    ext_ids = plan.ext_ids(rank)
    dinv_ext = dinv_full[ext_ids].contiguous()
    if hub_mb:
        plan.hub_rows = B.mark_hubs(plan.nbr_all, plan.n_ext, w.in_dim * 4, hub_mb << 20)
        assert 0 < plan.hub_rows <= (hub_mb << 20) // (w.in_dim * 4)
        marked = (plan.nbr_all.cpu().numpy() < 0)
        assert marked.any() and not marked.all()
    P = model.named_parameter_arrays()
    Wm, bm = P["gnn_convs_0_conv_lin_weight"], P["gnn_convs_0_conv_bias"]
    x_ext = B.to_device(x)[ext_ids].contiguous()          # what the exchange would have delivered
    y = B.empty((part.n_local, Wm.shape[0]))
    args = (x_ext, y, plan, dinv_ext, B.to_device(Wm), B.to_device(bm), None, 1)
    B.gcn_layer_halo(*args, 1, w.in_dim, Wm.shape[0])
    B.gcn_layer_halo(*args, 2, w.in_dim, Wm.shape[0])
    torch.cuda.synchronize()
    two_phase = y.cpu().numpy().copy()
    ref = _gcn_rows_ref(orc, x, coo, n, Wm, bm)[r0:r1]
    assert rel_err(two_phase, ref) < TOL
    y.zero_()
    B.gcn_layer_halo(*args, 3, w.in_dim, Wm.shape[0])
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), two_phase)     # phase 3 == phase 1 then phase 2, bit for bit


def test_halo_pack_signal_wait(setup):
    import ctypes as C
    import torch

    gnnb, *_ = setup
    from gnn_builder_b200 import _lib
    from gnn_builder_b200.distributed import CudaBackend

    lib = _lib.load()
    B = CudaBackend()
    rng = np.random.default_rng(5)
    for F in (128, 20, 7):
        xs = torch.from_numpy(rng.uniform(-1, 1, (1000, F)).astype(np.float32)).cuda()
        counts = [0, 300, 0, 555]
        idx = torch.from_numpy(rng.integers(0, 1000, sum(counts)).astype(np.int32)).cuda()

        class Plan:
            send_counts = counts
            send_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            send_idx = idx

        send = torch.zeros((sum(counts), F), device="cuda")
        B.pack_rows(xs, F, Plan, send)
        torch.cuda.synchronize()
        assert torch.equal(send, xs[idx.long()])
    # flags: signal then wait returns at once; waiting for an epoch nobody signals times out
    # (~2 s of SM clocks) and reports it instead of hanging
    flags_t, flags_ptr, handle = B.ipc_alloc(8 * 32)
    assert len(handle) == 64 and int(flags_t.sum().item()) == 0
    timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    B.halo_signal([flags_ptr + 8 * p for p in range(4)], 3)
    B.halo_wait(flags_ptr, 4, 3, timed_out)
    torch.cuda.synchronize()
    assert int(timed_out.item()) == 0
    assert flags_t.view(torch.int64)[:4].tolist() == [3, 3, 3, 3]
    B.halo_wait(flags_ptr, 4, 4, timed_out)
    torch.cuda.synchronize()
    assert int(timed_out.item()) == 1
    B.release()


def test_large_graph_hub_hints_do_not_change_results(setup, orc, monkeypatch):
    """the layerwise path marks hub sources (L2 evict_last) when the feature matrix exceeds L2;
    the marking only changes cache policy: results are bit-identical with it forced on and off"""
    import torch

    gnnb, w, model, n, x, coo = setup
    from gnn_builder_b200 import _lib
    from gnn_builder_b200.distributed import CudaBackend

    B = CudaBackend()
    nbr = torch.from_numpy(np.random.default_rng(1).zipf(1.6, 200000).clip(1, 5000).astype(np.int32) - 1).cuda()
    ref_cnt = np.bincount(nbr.cpu().numpy(), minlength=5000)
    marked_tbl = nbr.clone()
    n_hubs = B.mark_hubs(marked_tbl, 5000, 512, 100 * 512)       # room for 100 rows
    assert 0 < n_hubs <= 100
    got = marked_tbl.cpu().numpy()
    assert np.array_equal(got & 0x7fffffff, nbr.cpu().numpy())
    hubs = np.unique(got[got < 0] & 0x7fffffff)
    assert hubs.size == n_hubs
    thr = ref_cnt[hubs].min()
    assert (ref_cnt[np.setdiff1d(np.arange(5000), hubs)] < thr).all()   # exactly the top sources
