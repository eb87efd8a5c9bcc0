"""Multi-GPU execution on one 8 x B200 box, one process per GPU (torch.distributed, NCCL).

Two ways the hot path shards:

* **independent graphs** (BASELINE configs 1-4): a batch is cut into contiguous ranges of graphs
  balanced by node count; every rank runs its shard through its own ``Engine``; there is NO
  data-path collective (``run_sharded`` optionally gathers the per-graph outputs at the end).
* **one large graph** (BASELINE config 5, GCN): 1D partition by destination row.  Rank g owns
  rows ``[g*n/W, (g+1)*n/W)``, their CSR slice (in-edges, global source ids) and the matching
  feature rows.  Per layer the feature shards are all-gathered over NVLink (for a power-law
  graph the halo is practically every remote row, so the halo exchange is ``all_gather`` of the
  ``[n/W][F]`` shards), then each rank aggregates + transforms its own rows.  The in-degree
  table is all-gathered once.  Pooling: local partials + ``all_reduce`` (sum / max).

The collective plumbing is backend-agnostic (``gloo`` on CPU for the host-logic tests); the
compute goes through a small backend object -- ``CudaBackend`` (the C-ABI library) by default.
There is no CPU compute path in this package: the gloo tests inject a checker backend.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .data import GraphBatch


# ------------------------------------------------------------------ independent graphs
def shard_ranges(node_ptr: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous graph ranges [g0, g1) per rank, balanced by total node count."""
    n_graphs = int(node_ptr.shape[0] - 1)
    total = int(node_ptr[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        g = int(np.searchsorted(node_ptr, target, side="left"))
        g = min(max(g, bounds[-1]), n_graphs)
        bounds.append(g)
    bounds.append(n_graphs)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_batch(batch: GraphBatch, rank: int, world: int) -> Tuple[GraphBatch, int, int]:
    g0, g1 = shard_ranges(batch.node_ptr, world)[rank]
    return batch.slice(g0, g1), g0, g1


def run_sharded(engine, batch: GraphBatch, rank: int, world: int, gather: bool = False,
                dist=None) -> np.ndarray:
    """Run this rank's shard; with ``gather`` every rank returns the full [G][out] array."""
    shard, g0, g1 = shard_batch(batch, rank, world)
    out_local = engine.run(shard) if shard.n_graphs else np.zeros((0, engine.out_dim), np.float32)
    if not gather or world == 1:
        return out_local
    import torch

    ranges = shard_ranges(batch.node_ptr, world)
    cap = max(b - a for a, b in ranges)  # all_gather needs equal shapes: pad, gather, trim
    mine = torch.zeros((cap, engine.out_dim), dtype=torch.float32)
    mine[: out_local.shape[0]] = torch.from_numpy(np.ascontiguousarray(out_local))
    parts = [torch.empty((cap, engine.out_dim), dtype=torch.float32) for _ in ranges]
    dist.all_gather(parts, mine)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, ranges)], 0).numpy()


# ------------------------------------------------------------------ one large graph
@dataclass
class RowPartition:
    n_total: int
    world: int

    def __post_init__(self):
        if self.n_total % self.world != 0:
            raise ValueError("row partition needs num_nodes divisible by world size (pad the graph)")
        self.n_local = self.n_total // self.world

    def rows(self, rank: int) -> Tuple[int, int]:
        return rank * self.n_local, (rank + 1) * self.n_local

    def owner(self, node_ids: np.ndarray) -> np.ndarray:
        return node_ids // self.n_local

    def local_edges(self, coo: np.ndarray, rank: int) -> np.ndarray:
        """In-edges of the owned rows, COO order preserved (keeps the neighbor order stable)."""
        r0, r1 = self.rows(rank)
        m = (coo[:, 1] >= r0) & (coo[:, 1] < r1)
        return np.ascontiguousarray(coo[m])


class CudaBackend:
    """Compute through libgnnb_b200.so on torch CUDA tensors (current stream)."""

    def __init__(self):
        import torch

        from . import _lib

        self.torch = torch
        self._lib = _lib
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _p(self, t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def to_device(self, a: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def empty(self, shape, dtype="float32"):
        return self.torch.empty(shape, dtype=getattr(self.torch, dtype), device=self.device)

    def partition_tables(self, coo_local, row_begin, n_local):
        e = int(coo_local.shape[0])
        ind, off = self.empty((n_local,), "int32"), self.empty((n_local,), "int32")
        nbr = self.empty((max(e, 1),), "int32")
        self._lib.check(self.lib.gnnb_partition_tables(self._p(coo_local), row_begin, n_local, e,
                                                       self._p(ind), self._p(off), self._p(nbr),
                                                       self._stream()))
        return ind, off, nbr

    def dinv(self, in_deg_full):
        out = self.empty((in_deg_full.shape[0],))
        self._lib.check(self.lib.gnnb_degree_inv_sqrt(self._p(in_deg_full), self._p(out),
                                                      int(in_deg_full.shape[0]), self._stream()))
        return out

    def gcn_layer(self, x_full, tables, dinv_full, row_begin, W, b, skip, act):
        ind, off, nbr = tables
        n_local, n_total = int(ind.shape[0]), int(x_full.shape[0])
        fo, fi = int(W.shape[0]), int(W.shape[1])
        y = self.empty((n_local, fo))
        self._lib.check(self.lib.gnnb_gcn_conv_partition(
            n_local, row_begin, n_total, int(nbr.shape[0]), self._p(x_full), self._p(y),
            self._p(off), self._p(nbr), self._p(ind), self._p(dinv_full), self._p(W), self._p(b),
            self._p(skip), fi, fo, act, self._stream()))
        return y

    def pool_partial(self, x_local):
        n, f = int(x_local.shape[0]), int(x_local.shape[1])
        out = self.empty((2, f))
        self._lib.check(self.lib.gnnb_pool_partial(self._p(x_local.contiguous()), n, f, self._p(out),
                                                   self._stream()))
        return out[0], out[1]

    def head(self, pooled, linears, mlp_act, out_act):
        from . import layers

        h = pooled
        for j, (W, b) in enumerate(linears):
            h = layers.linear(h, W, b)
            last = j == len(linears) - 1
            a = out_act if last else mlp_act
            if a:
                h = layers.apply_activation(a, h)
        return h

    def all_gather_rows(self, dist, x_local, n_total):
        full = self.empty((n_total, x_local.shape[1]), str(x_local.dtype).split(".")[-1])
        dist.all_gather_into_tensor(full, x_local.contiguous())
        return full

    def all_reduce(self, dist, t, op):
        dist.all_reduce(t, op=op)
        return t


class LargeGraphGCN:
    """GCN GNNModel over one large row-partitioned graph (BASELINE config 5)."""

    def __init__(self, model, n_total: int, rank: int, world: int, dist=None, backend=None):
        d = model.describe()
        if d["conv_type"] != 0:
            raise NotImplementedError("the row-partitioned path implements GCN (BASELINE config 5)")
        self.model, self.desc = model, d
        self.part = RowPartition(n_total, world)
        self.rank, self.world, self.dist = rank, world, dist
        self.backend = backend if backend is not None else CudaBackend()
        params = model.named_parameter_arrays()
        names = list(params)
        nh = d["mlp_num_linear"]
        B = self.backend
        self.head = [(B.to_device(params[names[2 * j]]), B.to_device(params[names[2 * j + 1]]))
                     for j in range(nh)]
        self.layers = []
        for k in range(d["num_layers"]):  # [conv_bias, conv_lin_weight] per layer
            b, W = params[names[2 * nh + 2 * k]], params[names[2 * nh + 2 * k + 1]]
            self.layers.append((B.to_device(W), B.to_device(b)))
        self.tables = None
        self.dinv_full = None

    def setup(self, coo_local: np.ndarray):
        """Build this rank's CSR slice and the global 1/sqrt(1+deg) table (one all-gather)."""
        B = self.backend
        r0, _ = self.part.rows(self.rank)
        self.tables = B.partition_tables(B.to_device(coo_local.astype(np.int32)), r0,
                                         self.part.n_local)
        ind_local = self.tables[0]
        if self.world > 1:
            ind_full = B.all_gather_rows(self.dist, ind_local.view(-1, 1), self.part.n_total).view(-1)
        else:
            ind_full = ind_local
        self.dinv_full = B.dinv(ind_full)
        return self

    def forward(self, x_local, return_embeddings: bool = False):
        """x_local: this rank's feature rows (backend tensor or numpy).  Returns the model output
        (identical on every rank) and optionally the rank's node-embedding rows."""
        B, d = self.backend, self.desc
        if isinstance(x_local, np.ndarray):
            x_local = B.to_device(x_local.astype(np.float32))
        r0, _ = self.part.rows(self.rank)
        L = d["num_layers"]
        for k, (W, b) in enumerate(self.layers):
            x_full = (B.all_gather_rows(self.dist, x_local, self.part.n_total)
                      if self.world > 1 else x_local)
            do_skip = bool(d["skip"]) and k != 0 and k != L - 1
            x_local = B.gcn_layer(x_full, self.tables, self.dinv_full, r0, W, b,
                                  x_local if do_skip else None, d["gnn_act"])
        s, mx = B.pool_partial(x_local)
        if self.world > 1:
            s = B.all_reduce(self.dist, s, self.dist.ReduceOp.SUM)
            mx = B.all_reduce(self.dist, mx, self.dist.ReduceOp.MAX)
        pools = []
        for pid in d["pools"]:
            pools.append({0: s, 1: s / float(self.part.n_total), 2: mx}[pid])
        pooled = _cat(pools)
        out = B.head(pooled, self.head, d["mlp_act"], d["out_act"])
        return (out, x_local) if return_embeddings else out


def _cat(ts):
    if hasattr(ts[0], "device"):
        import torch

        return torch.cat([t.reshape(-1) for t in ts])
    return np.concatenate([np.asarray(t).reshape(-1) for t in ts])
