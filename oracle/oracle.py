"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the checker.

``Oracle``   : the C restatement (oracle/gnnb_oracle.c -> oracle/_ref/libgnnb_oracle.so).
``RefLayers``: the reference's own templates (oracle/ref_layers.cpp -> libgnnb_ref_layers.so).
``RefModel`` : a ``<name>_top`` rendered by the reference's own code generator and compiled
               (oracle/_ref/models/<name>/lib<name>.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
i64p = C.POINTER(C.c_longlong)


def _f(a):
    return a.ctypes.data_as(f32p)


def _i(a):
    return a.ctypes.data_as(i32p)


def _c32(a):
    return np.ascontiguousarray(a, np.float32)


def _ci(a):
    return np.ascontiguousarray(a, np.int32)


class OrcModelDesc(C.Structure):
    _fields_ = [
        ("conv_type", C.c_int), ("num_layers", C.c_int), ("in_dim", C.c_int),
        ("hidden_dim", C.c_int), ("out_dim", C.c_int), ("skip", C.c_int), ("gnn_act", C.c_int),
        ("gin_eps", C.c_float), ("pna_delta", C.c_float), ("num_pools", C.c_int),
        ("pools", C.c_int * 4), ("mlp_num_linear", C.c_int), ("mlp_hidden", C.c_int),
        ("mlp_out", C.c_int), ("mlp_act", C.c_int), ("out_act", C.c_int),
        ("gnn_p_in", C.c_int), ("gnn_p_hidden", C.c_int), ("gnn_p_out", C.c_int),
        ("mlp_p_in", C.c_int), ("mlp_p_hidden", C.c_int), ("mlp_p_out", C.c_int),
    ]


def desc_from_dict(d: dict) -> OrcModelDesc:
    """``d`` is ``GNNModel.describe()`` (gnn_builder_b200/models.py)."""
    m = OrcModelDesc()
    for k in ("conv_type", "num_layers", "in_dim", "hidden_dim", "out_dim", "skip", "gnn_act",
              "mlp_num_linear", "mlp_hidden", "mlp_out", "mlp_act", "out_act"):
        setattr(m, k, int(d[k]))
    m.gin_eps = float(d["gin_eps"])
    m.pna_delta = float(d["pna_delta"])
    m.num_pools = len(d["pools"])
    for i, p in enumerate(d["pools"]):
        m.pools[i] = int(p)
    for k in ("gnn_p_in", "gnn_p_hidden", "gnn_p_out", "mlp_p_in", "mlp_p_hidden", "mlp_p_out"):
        setattr(m, k, int(d.get(k, 1)))
    return m


def ensure_built():
    sys.path.insert(0, str(HERE))
    import build_ref

    return build_ref.build_oracle()


class Oracle:
    def __init__(self):
        so = OUT / "libgnnb_oracle.so"
        if not so.exists():
            ensure_built()
        self.lib = L = C.CDLL(str(so))
        L.orc_activation.restype = C.c_float
        L.orc_activation.argtypes = [C.c_int, C.c_float]
        L.orc_model_forward.restype = C.c_int
        L.orc_model_forward_batch.restype = C.c_int
        L.orc_model_num_params.restype = C.c_int

    # ---- primitives
    def activation(self, act: int, x):
        x = _c32(x)
        y = np.empty_like(x)
        self.lib.orc_apply_activation(C.c_int(act), _f(x), _f(y), C.c_int(x.size))
        return y

    def linear(self, x, W, b, block_in: int = 1):
        x, W = _c32(x), _c32(W)
        out_size, in_size = W.shape
        b = _c32(b) if b is not None else None
        y = np.empty(out_size, np.float32)
        self.lib.orc_linear(_f(x), _f(y), _f(W), _f(b) if b is not None else None,
                            C.c_int(in_size), C.c_int(out_size), C.c_int(block_in))
        return y

    def degree_tables(self, coo, n):
        coo = _ci(coo).reshape(-1, 2)
        ind, outd = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.lib.orc_compute_degree_tables(_i(coo), _i(ind), _i(outd), C.c_int(n),
                                           C.c_int(coo.shape[0]))
        return ind, outd

    def neighbor_tables(self, coo, in_deg, with_edge_index=False):
        coo = _ci(coo).reshape(-1, 2)
        in_deg = _ci(in_deg)
        n, e = in_deg.shape[0], coo.shape[0]
        off, nbr = np.zeros(n, np.int32), np.zeros(e, np.int32)
        eidx = np.zeros(e, np.int32)
        self.lib.orc_compute_neighbor_and_edge_index_tables(
            _i(coo), _i(in_deg), _i(off), _i(nbr), _i(eidx), C.c_int(n), C.c_int(e))
        return (off, nbr, eidx) if with_edge_index else (off, nbr)

    def tables(self, coo, n):
        ind, outd = self.degree_tables(coo, n)
        off, nbr = self.neighbor_tables(coo, ind)
        return ind, outd, off, nbr

    # ---- convs (tables given)
    def gcn_conv(self, x, off, nbr, ind, W, b, p_in=1):
        x, W, b = _c32(x), _c32(W), _c32(b)
        n, fi = x.shape
        fo = W.shape[0]
        y = np.empty((n, fo), np.float32)
        self.lib.orc_gcn_conv(C.c_int(n), _f(x), _f(y), _i(_ci(off)), _i(_ci(nbr)), _i(_ci(ind)),
                              _f(W), _f(b), C.c_int(fi), C.c_int(fo), C.c_int(p_in))
        return y

    def gin_conv(self, x, off, nbr, ind, W0, b0, W1, b1, eps, p_in=1):
        x, W0, b0, W1, b1 = map(_c32, (x, W0, b0, W1, b1))
        n, fi = x.shape
        hid, fo = W0.shape[0], W1.shape[0]
        y = np.empty((n, fo), np.float32)
        self.lib.orc_gin_conv(C.c_int(n), _f(x), _f(y), _i(_ci(off)), _i(_ci(nbr)), _i(_ci(ind)),
                              _f(W0), _f(b0), _f(W1), _f(b1), C.c_float(eps), C.c_int(fi),
                              C.c_int(hid), C.c_int(fo), C.c_int(p_in))
        return y

    def gine_conv(self, x, ef, off, nbr, eidx, ind, We, be, W0, b0, W1, b1, eps, p_in=1):
        x, ef, We, be, W0, b0, W1, b1 = map(_c32, (x, ef, We, be, W0, b0, W1, b1))
        n, fi = x.shape
        hid, fo, fe = W0.shape[0], W1.shape[0], ef.shape[1]
        y = np.empty((n, fo), np.float32)
        self.lib.orc_gine_conv(C.c_int(n), _f(x), _f(y), _f(ef), _i(_ci(off)), _i(_ci(nbr)),
                               _i(_ci(eidx)), _i(_ci(ind)), _f(We), _f(be), _f(W0), _f(b0), _f(W1),
                               _f(b1), C.c_float(eps), C.c_int(fi), C.c_int(hid), C.c_int(fo),
                               C.c_int(fe), C.c_int(p_in))
        return y

    def sage_conv(self, x, off, nbr, ind, Wl, bl, Wr, p_in=1):
        x, Wl, bl, Wr = map(_c32, (x, Wl, bl, Wr))
        n, fi = x.shape
        fo = Wl.shape[0]
        y = np.empty((n, fo), np.float32)
        self.lib.orc_sage_conv(C.c_int(n), _f(x), _f(y), _i(_ci(off)), _i(_ci(nbr)), _i(_ci(ind)),
                               _f(Wl), _f(bl), _f(Wr), C.c_int(fi), C.c_int(fo), C.c_int(p_in))
        return y

    def pna_conv(self, x, off, nbr, ind, Wpre, bpre, Wpost, bpost, Wlin, blin, delta, p_in=1,
                 p_out=1):
        x, Wpre, bpre, Wpost, bpost, Wlin, blin = map(_c32, (x, Wpre, bpre, Wpost, bpost, Wlin,
                                                             blin))
        n, fi = x.shape
        fo = Wpost.shape[0]
        y = np.empty((n, fo), np.float32)
        self.lib.orc_pna_conv(C.c_int(n), _f(x), _f(y), _i(_ci(off)), _i(_ci(nbr)), _i(_ci(ind)),
                              _f(Wpre), _f(bpre), _f(Wpost), _f(bpost), _f(Wlin), _f(blin),
                              C.c_float(delta), C.c_int(fi), C.c_int(fo), C.c_int(p_in),
                              C.c_int(p_out))
        return y

    def lg_conv(self, x, off, nbr, ind):
        x = _c32(x)
        y = np.empty_like(x)
        self.lib.orc_lg_conv(C.c_int(x.shape[0]), _f(x), _f(y), _i(_ci(off)), _i(_ci(nbr)),
                             _i(_ci(ind)), C.c_int(x.shape[1]))
        return y

    def simple_conv(self, x, off, nbr, ind):
        x = _c32(x)
        y = np.empty_like(x)
        self.lib.orc_simple_conv(C.c_int(x.shape[0]), _f(x), _f(y), _i(_ci(off)), _i(_ci(nbr)),
                                 _i(_ci(ind)), C.c_int(x.shape[1]))
        return y

    def pool(self, kind: str, x):
        x = _c32(x)
        n, f = x.shape
        out = np.empty(f, np.float32)
        getattr(self.lib, f"orc_global_{kind}_pool")(C.c_int(n), _f(x), C.c_int(f), _f(out))
        return out

    # ---- whole model
    def _param_array(self, desc, params):
        params = [_c32(p) for p in params]
        n = self.lib.orc_model_num_params(C.byref(desc))
        assert n == len(params), f"oracle expects {n} parameter arrays, got {len(params)}"
        arr = (f32p * n)(*[_f(p) for p in params])
        return arr, params

    def model_forward(self, desc_dict, params, x, coo, return_node_emb=False):
        """params: list of arrays in the reference's flat order (mlp_head first)."""
        d = desc_from_dict(desc_dict)
        arr, keep = self._param_array(d, params)
        x, coo = _c32(x), _ci(coo).reshape(-1, 2)
        n, e = x.shape[0], coo.shape[0]
        out = np.empty(d.mlp_out, np.float32)
        emb_dim = d.out_dim if d.num_layers > 0 else d.in_dim
        emb = np.empty((n, emb_dim), np.float32)
        rc = self.lib.orc_model_forward(C.byref(d), arr, _f(x), _i(coo), C.c_int(n), C.c_int(e),
                                        _f(out), _f(emb))
        assert rc == 0, rc
        return (out, emb) if return_node_emb else out

    def model_forward_batch(self, desc_dict, params, batch):
        d = desc_from_dict(desc_dict)
        arr, keep = self._param_array(d, params)
        x, coo = _c32(batch.x), _ci(batch.coo)
        nptr = np.ascontiguousarray(batch.node_ptr, np.int64)
        eptr = np.ascontiguousarray(batch.edge_ptr, np.int64)
        out = np.empty((batch.n_graphs, d.mlp_out), np.float32)
        rc = self.lib.orc_model_forward_batch(C.byref(d), arr, _f(x), _i(coo),
                                              nptr.ctypes.data_as(i64p), eptr.ctypes.data_as(i64p),
                                              C.c_int(batch.n_graphs), _f(out))
        assert rc == 0, rc
        return out


# --------------------------------------------------------------------------- the real reference

def ref_available() -> bool:
    return (OUT / "libgnnb_ref_layers.so").exists()


class RefLayers:
    """extern "C" instantiations of the reference's own templates (oracle/ref_layers.cpp)."""

    def __init__(self):
        self.lib = C.CDLL(str(OUT / "libgnnb_ref_layers.so"))
        self.lib.ref_activation.restype = C.c_float
        self.lib.ref_activation.argtypes = [C.c_int, C.c_float]
        self.max_nodes = self.lib.ref_max_nodes()
        self.max_edges = self.lib.ref_max_edges()

    def has(self, name: str) -> bool:
        return hasattr(self.lib, name)

    def activation(self, act, x):
        return np.asarray([self.lib.ref_activation(act, float(v)) for v in np.ravel(x)],
                          np.float32).reshape(np.shape(x))

    def tables(self, coo, n, with_edge_index=False):
        coo = _ci(coo).reshape(-1, 2)
        e = coo.shape[0]
        ind, outd = np.zeros(n, np.int32), np.zeros(n, np.int32)
        off, nbr, eidx = np.zeros(n, np.int32), np.zeros(e, np.int32), np.zeros(e, np.int32)
        self.lib.ref_compute_degree_tables(_i(coo), _i(ind), _i(outd), C.c_int(n), C.c_int(e))
        self.lib.ref_compute_neighbor_and_edge_index_tables(
            _i(coo), _i(ind), _i(outd), _i(off), _i(nbr), _i(eidx), C.c_int(n), C.c_int(e))
        if with_edge_index:
            return ind, outd, off, nbr, eidx
        return ind, outd, off, nbr

    def conv(self, kind, x, coo, tables, weights, scalar=None, fo=None):
        """kind in gcn|gin|sage|pna; weights in the reference's argument order."""
        ind, outd, off, nbr = tables
        x, coo = _c32(x), _ci(coo).reshape(-1, 2)
        n, fi = x.shape
        fn = getattr(self.lib, f"ref_{kind}_conv_{fi}_{fo}")
        y = np.zeros((n, fo), np.float32)
        ws = [_c32(w) for w in weights]
        args = [C.c_int(n), C.c_int(coo.shape[0]), _f(x), _f(y), _i(coo), _i(_ci(off)),
                _i(_ci(nbr)), _i(_ci(ind)), _i(_ci(outd))] + [_f(w) for w in ws]
        if scalar is not None:
            args.append(C.c_float(scalar))
        fn(*args)
        return y

    def same_conv(self, kind, x, coo, tables):
        ind, outd, off, nbr = tables
        x, coo = _c32(x), _ci(coo).reshape(-1, 2)
        n, f = x.shape
        y = np.zeros((n, f), np.float32)
        getattr(self.lib, f"ref_{kind}_conv_{f}")(
            C.c_int(n), C.c_int(coo.shape[0]), _f(x), _f(y), _i(coo), _i(_ci(off)), _i(_ci(nbr)),
            _i(_ci(ind)), _i(_ci(outd)))
        return y

    def gine_conv(self, x, ef, coo, tables5, weights, eps):
        ind, outd, off, nbr, eidx = tables5
        x, ef, coo = _c32(x), _c32(ef), _ci(coo).reshape(-1, 2)
        n = x.shape[0]
        y = np.zeros((n, 8), np.float32)
        ws = [_c32(w) for w in weights]
        self.lib.ref_gine_conv_8_8_16(
            C.c_int(n), C.c_int(coo.shape[0]), _f(x), _f(y), _f(ef), _i(coo), _i(_ci(off)),
            _i(_ci(nbr)), _i(_ci(eidx)), _i(_ci(ind)), _i(_ci(outd)), *[_f(w) for w in ws],
            C.c_float(eps))
        return y

    def pool(self, kind, x):
        x = _c32(x)
        n, f = x.shape
        out = np.zeros(f, np.float32)
        getattr(self.lib, f"ref_global_{kind}_pool_{f}")(C.c_int(n), _f(x), _f(out))
        return out

    def linear(self, x, W, b, buffered=False):
        x, W, b = _c32(x), _c32(W), _c32(b)
        fo, fi = W.shape
        y = np.zeros(fo, np.float32)
        name = f"ref_linear{'_buffered' if buffered else ''}_{fi}_{fo}"
        getattr(self.lib, name)(_f(x), _f(y), _f(W), _f(b))
        return y


class RefModel:
    """``<name>_top`` compiled from the reference's own generated model.cpp (one graph per call,
    not re-entrant: file-scope static buffers, model.cpp.jinja:7-22,197-209)."""

    def __init__(self, name: str):
        d = OUT / "models" / name
        self.dir = d
        self.manifest = json.loads((d / "manifest.json").read_text())
        self.lib = C.CDLL(str(d / f"lib{name}.so"))
        self.top = getattr(self.lib, f"{name}_top")
        self.top.restype = None
        m = self.manifest
        self.max_nodes, self.max_edges = m["max_nodes"], m["max_edges"]
        self.in_dim, self.out_dim = m["in_dim"], m["out_dim"]
        self._x = np.zeros((self.max_nodes, self.in_dim), np.float32)
        self._coo = np.zeros((self.max_edges, 2), np.int32)
        self._out = np.zeros(self.out_dim, np.float32)
        self._params = None
        self._loaded = False

    @property
    def param_names(self):
        return self.manifest["param_names"]

    def set_params(self, params: dict):
        self._params = []
        for name, shape in zip(self.manifest["param_names"], self.manifest["param_shapes"]):
            a = _c32(params[name])
            assert list(a.shape) == list(shape), (name, a.shape, shape)
            self._params.append(a)
        self._loaded = False

    def __call__(self, x, coo):
        x, coo = _c32(x), _ci(coo).reshape(-1, 2)
        n, e = x.shape[0], coo.shape[0]
        assert n <= self.max_nodes and e <= self.max_edges
        self._x[:n] = x
        self._coo[:e] = coo
        flag = 0 if self._loaded else 1
        self.top(_f(self._x), _i(self._coo), _f(self._out), C.c_int(n), C.c_int(e), C.c_int(flag),
                 *[_f(p) for p in self._params])
        self._loaded = True
        return self._out.copy()

    def run_batch(self, batch):
        out = np.empty((batch.n_graphs, self.out_dim), np.float32)
        for g in range(batch.n_graphs):
            x, coo = batch.graph(g)
            out[g] = self(x, coo)
        return out


def ref_big_gcn_rate(x, coo, n, W, b, num_rows: int):
    """Reference CPU baseline for the large-graph config: the reference's own
    compute_degree_tables / compute_neighbor_tables / gcn_conv<2000000, 40000000, 128, 128>
    driven on heap buffers.  Times one gcn_conv layer over the first ``num_rows`` destination rows
    of the graph.  Returns (edges_per_second, seconds, edges_done, table_seconds).  The reference
    keeps `int neighbors[MAX_NODES]` (8 MB) on the stack, hence the big-stack thread."""
    import threading
    import time

    lib = C.CDLL(str(OUT / "libgnnb_ref_layers.so"))
    x, W, b = _c32(x), _c32(W), _c32(b)
    coo = _ci(coo).reshape(-1, 2)
    e = coo.shape[0]
    ind, outd = np.zeros(n, np.int32), np.zeros(n, np.int32)
    off, nbr = np.zeros(n, np.int32), np.zeros(e, np.int32)
    y = np.zeros((max(num_rows, 1), 128), np.float32)
    res = {}

    def work():
        t0 = time.perf_counter()
        lib.ref_big_tables(_i(coo), _i(ind), _i(outd), _i(off), _i(nbr), C.c_int(n), C.c_int(e))
        t1 = time.perf_counter()
        lib.ref_big_gcn_conv_128_128(C.c_int(num_rows), C.c_int(e), _f(x), _f(y), _i(coo), _i(off),
                                     _i(nbr), _i(ind), _i(outd), _f(W), _f(b))
        res["t_tables"], res["t_conv"] = t1 - t0, time.perf_counter() - t1

    threading.stack_size(512 * 1024 * 1024)
    th = threading.Thread(target=work)
    th.start()
    th.join()
    threading.stack_size(0)
    edges_done = int(ind[:num_rows].sum())
    return edges_done / res["t_conv"], res["t_conv"], edges_done, res["t_tables"], y
