"""Host-side logic of the N > 1 paths on CPU: world_size-2 `gloo` process groups.

The product has no CPU compute path, so these tests inject a *checker* backend (numpy + the
oracle) into the orchestration code and verify what the orchestration is responsible for: shard
ranges, the row partition, the all-gather / all-reduce plumbing, the halo plan (ext index space, split CSR, all-to-all of the request lists and of
the feature rows) and the assembly of results."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


class NumpyBackend:
    """Checker backend for LargeGraphGCN: numpy restatement of the partitioned GCN layer."""

    def to_device(self, a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a))

    def partition_tables(self, coo_local, row_begin, n_local):
        import torch
        coo = coo_local.numpy()
        dst = coo[:, 1] - row_begin
        order = np.argsort(dst, kind="stable")
        ind = np.bincount(dst, minlength=n_local).astype(np.int32)
        off = np.concatenate([[0], np.cumsum(ind)[:-1]]).astype(np.int32)
        return torch.from_numpy(ind), torch.from_numpy(off), torch.from_numpy(coo[order, 0].copy())

    def dinv(self, in_deg_full):
        return 1.0 / (1.0 + in_deg_full.float()).sqrt()

    # the exchange runs "on another stream" in the product; here everything is synchronous
    def fork(self):
        pass

    def join(self):
        pass

    def comm_ctx(self):
        import contextlib
        return contextlib.nullcontext()

    def empty(self, shape, dtype="float32"):
        import torch
        return torch.zeros(shape, dtype=getattr(torch, dtype))

    def pack_rows(self, x_own, F, plan, send):
        send[:] = x_own.view(-1, F)[plan.send_idx.long()]

    def gcn_layer_halo(self, x_ext, y_local, plan, dinv_ext, W, b, skip, act, phase, fi, fo,
                       row_begin=0, row_count=0):
        assert row_count in (0, plan.n_local), "the checker backend runs whole layers (nccl transport)"
        """numpy/torch restatement of gnnb_gcn_conv_halo on the ext index space"""
        import torch
        n_local = plan.n_local
        x = x_ext.view(-1, fi)
        if phase & 1:
            self._agg = torch.zeros(n_local, fi)
            parts = [(plan.own_cnt, plan.own_nbr)]
        else:
            parts = []
        if phase & 2:
            parts.append((plan.halo_cnt, plan.halo_nbr))
        for cnt, nbr in parts:
            rows = torch.repeat_interleave(torch.arange(n_local), cnt.long())
            src = (nbr.long() & 0x7fffffff)
            self._agg.index_add_(0, rows, x[src] * dinv_ext[src].view(-1, 1))
        if not (phase & 2):
            return
        dv = dinv_ext[:n_local].view(-1, 1)
        agg = self._agg * dv + x[:n_local] * dv * dv
        y = agg @ W.T + b
        if skip is not None:
            y = y + skip.view(n_local, fi)
        y_local.view(n_local, fo)[:] = torch.relu(y) if act == 1 else y

    def pool_partial(self, x_local):
        return x_local.sum(0), x_local.max(0).values

    def head(self, pooled, linears, mlp_act, out_act):
        import torch
        h = pooled
        for j, (W, b) in enumerate(linears):
            h = W @ h + b
            if j < len(linears) - 1 and mlp_act == 1:
                h = torch.relu(h)
        return h

    def all_gather_rows(self, dist, x_local, n_total):
        import torch
        full = torch.empty((n_total, x_local.shape[1]), dtype=x_local.dtype)
        dist.all_gather_into_tensor(full, x_local.contiguous())
        return full

    def all_reduce(self, dist, t, op):
        dist.all_reduce(t, op=op)
        return t


class OracleEngine:
    """Checker stand-in for Engine (tests only): runs a shard through the CPU oracle."""

    def __init__(self, model):
        sys.path.insert(0, str(ROOT / "oracle"))
        from oracle import Oracle
        self.orc, self.model = Oracle(), model
        self.params = list(model.named_parameter_arrays().values())
        self.out_dim = model.output_features_dim

    def run(self, batch):
        return self.orc.model_forward_batch(self.model.describe(), self.params, batch)


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dataclasses
        import gnn_builder_b200 as gnnb
        from gnn_builder_b200 import distributed as D
        from gnn_builder_b200.configs import C1, C5
        from gnn_builder_b200.data import make_powerlaw_graph
        from oracle import Oracle

        # ---- independent graphs: shard, run, gather
        w = dataclasses.replace(C1, hidden_dim=12, in_dim=5, mlp_hidden_dim=8)
        model = gnnb.build_model(w, seed=0)
        batch = gnnb.make_molecular_batch(37, w.mu_nodes, w.mu_edges, w.in_dim, seed=3)
        eng = OracleEngine(model)
        out = D.run_sharded(eng, batch, rank, world, gather=True, dist=dist)
        ref = eng.run(batch)
        assert np.array_equal(out, ref), "sharded + gathered outputs differ from the single run"

        # ---- one large graph: row partition + all-gather per layer
        n = 600
        wl = dataclasses.replace(C5, in_dim=16, hidden_dim=16, out_dim=4, mlp_hidden_dim=8)
        big = gnnb.build_model(wl, seed=1)
        x, coo = make_powerlaw_graph(n, 6, wl.in_dim, seed=9, max_degree=80)
        part = D.RowPartition(n, world)
        r0, r1 = part.rows(rank)
        runner = D.LargeGraphGCN(big, n, rank, world, dist=dist, backend=NumpyBackend())
        runner.setup(part.local_edges(coo, rank))
        out, emb_local = runner.forward(x[r0:r1], return_embeddings=True)
        orc = Oracle()
        ref_out, ref_emb = orc.model_forward(big.describe(),
                                             list(big.named_parameter_arrays().values()), x, coo,
                                             return_node_emb=True)
        assert np.abs(emb_local.numpy() - ref_emb[r0:r1]).max() < 1e-5
        assert np.abs(out.numpy() - ref_out).max() < 1e-4 * max(1.0, np.abs(ref_out).max())
        # the halo plan: ext space = owned rows + ONE copy of each referenced remote row
        plan = runner.plan
        refs = np.unique(part.local_edges(coo, rank)[:, 0])
        remote = refs[(refs < r0) | (refs >= r1)]
        assert np.array_equal(plan.halo_ids.numpy(), remote)
        assert plan.n_ext == (r1 - r0) + remote.size and runner.transport == "nccl"
        assert int(plan.own_cnt.sum() + plan.halo_cnt.sum()) == part.local_edges(coo, rank).shape[0]
        # a second forward reuses the buffers (odd/even ext buffers, send buffer) unchanged
        out2 = runner.forward(x[r0:r1])
        assert np.array_equal(out2.numpy(), out.numpy())
        Path(result_dir, f"ok_{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_world_size_n_gloo(tmp_path, world):
    import torch.multiprocessing as mp

    port = 29500 + ((os.getpid() + 13 * world) % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok_{r}").exists() for r in range(world))


def test_shard_ranges_balance_and_cover():
    sys.path.insert(0, str(ROOT))
    import gnn_builder_b200 as gnnb
    from gnn_builder_b200.distributed import RowPartition, shard_ranges

    b = gnnb.make_molecular_batch(1000, 18, 38, 11, seed=2)
    for world in (1, 2, 4, 8):
        r = shard_ranges(b.node_ptr, world)
        assert r[0][0] == 0 and r[-1][1] == b.n_graphs
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        nodes = [int(b.node_ptr[g1] - b.node_ptr[g0]) for g0, g1 in r]
        assert max(nodes) - min(nodes) <= 2 * int(np.diff(b.node_ptr).max())
    # row partition: every edge belongs to exactly one rank, order preserved
    rng = np.random.default_rng(0)
    coo = rng.integers(0, 64, (500, 2)).astype(np.int32)
    part = RowPartition(64, 4)
    pieces = [part.local_edges(coo, r) for r in range(4)]
    assert sum(p.shape[0] for p in pieces) == 500
    for r, p in enumerate(pieces):
        lo, hi = part.rows(r)
        assert ((p[:, 1] >= lo) & (p[:, 1] < hi)).all()
    with pytest.raises(ValueError):
        RowPartition(65, 4)


def test_halo_plan_block_ranges():
    """the per-block send ranges of the pipelined halo push: every peer's list is cut at the block
    bounds, blocks are whole 128-row GEMM tiles, and together they cover each list exactly once"""
    sys.path.insert(0, str(ROOT))
    import torch

    from gnn_builder_b200.data import make_powerlaw_graph
    from gnn_builder_b200.distributed import HaloPlan, RowPartition

    n, world, rank = 4096, 4, 1
    _, coo = make_powerlaw_graph(n, 6, 4, seed=3, max_degree=200)
    part = RowPartition(n, world)
    loc = part.local_edges(coo, rank)
    r0, r1 = part.rows(rank)
    dst = loc[:, 1] - r0
    order = np.argsort(dst, kind="stable")
    ind = torch.from_numpy(np.bincount(dst, minlength=part.n_local).astype(np.int32))
    plan = HaloPlan(ind, torch.from_numpy(loc[order, 0].copy()), rank, part, dist=None)
    # pretend peers 0, 2, 3 asked for sorted subsets of the owned rows
    rng = np.random.default_rng(0)
    lists = [np.sort(rng.choice(part.n_local, size=k, replace=False)) for k in (300, 0, 517, 64)]
    plan.send_counts = [len(x) for x in lists]
    plan.send_off = np.concatenate([[0], np.cumsum(plan.send_counts)]).astype(np.int64)
    plan.send_idx = torch.from_numpy(np.concatenate(lists).astype(np.int32))
    for nb in (1, 3, 4, 16):
        bounds, lo = plan.block_ranges(nb)
        assert bounds[0] == 0 and bounds[-1] == part.n_local and (np.diff(bounds) >= 0).all()
        assert all(b % 128 == 0 for b in bounds[1:-1])
        for p, lst in enumerate(lists):
            assert lo[0][p] == 0 and lo[-1][p] == len(lst)
            for b in range(nb):
                piece = lst[lo[b][p]: lo[b + 1][p]]
                assert ((piece >= bounds[b]) & (piece < bounds[b + 1])).all()
