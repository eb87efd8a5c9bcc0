"""The reference's layer-function API (gnn_builder_lib.h) as Python callables over the C-ABI.

Same names and argument meaning as the C++ templates; template ints become array shapes.  Arrays
may be numpy arrays (host; staged by the library) or torch CUDA tensors (used in place).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

FAST, STRICT = 0, 1


def _is_torch(a) -> bool:
    return hasattr(a, "data_ptr")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def _f32(a):
    return a if _is_torch(a) else np.ascontiguousarray(a, np.float32)


def _i32(a):
    return a if _is_torch(a) else np.ascontiguousarray(a, np.int32)


def _empty_like_space(ref, shape, dtype):
    if _is_torch(ref):
        import torch

        return torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), device=ref.device)
    return np.empty(shape, dtype)


def compute_degree_tables(edge_list, num_nodes: int):
    """lib:1051-1083 -> (in_degree_table, out_degree_table)"""
    coo = _i32(edge_list)
    e = int(coo.shape[0])
    ind = _empty_like_space(coo, (num_nodes,), np.int32)
    outd = _empty_like_space(coo, (num_nodes,), np.int32)
    _lib.check(_lib.load().gnnb_compute_degree_tables(_ptr(coo), _ptr(ind), _ptr(outd), num_nodes, e))
    return ind, outd


def compute_neighbor_tables(edge_list, in_degree_table, out_degree_table=None,
                            with_edge_index: bool = False):
    """lib:1086-1166 -> (neighbor_table_offsets, neighbor_table[, edge_index_table])"""
    coo, ind = _i32(edge_list), _i32(in_degree_table)
    n, e = int(ind.shape[0]), int(coo.shape[0])
    off = _empty_like_space(coo, (n,), np.int32)
    nbr = _empty_like_space(coo, (e,), np.int32)
    eidx = _empty_like_space(coo, (e,), np.int32) if with_edge_index else None
    outd = _i32(out_degree_table) if out_degree_table is not None else None
    _lib.check(_lib.load().gnnb_compute_neighbor_and_edge_index_tables(
        _ptr(coo), _ptr(ind), _ptr(outd), _ptr(off), _ptr(nbr), _ptr(eidx), n, e))
    return (off, nbr, eidx) if with_edge_index else (off, nbr)


def linear(x, weight, bias, math: int = FAST):
    """lib:808-1003; x is [in] or [rows][in]"""
    x, weight, bias = _f32(x), _f32(weight), _f32(bias)
    single = x.ndim == 1
    rows = 1 if single else int(x.shape[0])
    out_size, in_size = int(weight.shape[0]), int(weight.shape[1])
    y = _empty_like_space(x, (out_size,) if single else (rows, out_size), np.float32)
    _lib.check(_lib.load().gnnb_linear(_ptr(x), _ptr(y), _ptr(weight), _ptr(bias), rows, in_size,
                                       out_size, math))
    return y


def apply_activation(act: int, x):
    x = _f32(x)
    y = _empty_like_space(x, tuple(x.shape), np.float32)
    n = int(np.prod(x.shape)) if len(x.shape) else 1
    _lib.check(_lib.load().gnnb_apply_activation(act, _ptr(x), _ptr(y), n))
    return y


def _conv_prologue(x, edge_list, offsets, nbr, in_deg, out_deg, f_out):
    x = _f32(x)
    n, e = int(x.shape[0]), int(nbr.shape[0])
    y = _empty_like_space(x, (n, f_out), np.float32)
    coo = _i32(edge_list) if edge_list is not None else None
    outd = _i32(out_deg) if out_deg is not None else None
    return x, y, n, e, coo, _i32(offsets), _i32(nbr), _i32(in_deg), outd


def gcn_conv(x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table,
             out_degree_table, weight, bias, math: int = FAST):
    """lib:1291-1387"""
    weight, bias = _f32(weight), _f32(bias)
    fo, fi = int(weight.shape[0]), int(weight.shape[1])
    x, y, n, e, coo, off, nbr, ind, outd = _conv_prologue(
        x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table, out_degree_table, fo)
    _lib.check(_lib.load().gnnb_gcn_conv(n, e, _ptr(x), _ptr(y), _ptr(coo), _ptr(off), _ptr(nbr),
                                         _ptr(ind), _ptr(outd), _ptr(weight), _ptr(bias), fi, fo,
                                         math))
    return y


def gin_conv(x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table,
             out_degree_table, mlp_0_weight, mlp_0_bias, mlp_1_weight, mlp_1_bias, gin_eps: float,
             math: int = FAST):
    """lib:1440-1549"""
    w0, b0, w1, b1 = map(_f32, (mlp_0_weight, mlp_0_bias, mlp_1_weight, mlp_1_bias))
    hid, fi, fo = int(w0.shape[0]), int(w0.shape[1]), int(w1.shape[0])
    x, y, n, e, coo, off, nbr, ind, outd = _conv_prologue(
        x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table, out_degree_table, fo)
    _lib.check(_lib.load().gnnb_gin_conv(n, e, _ptr(x), _ptr(y), _ptr(coo), _ptr(off), _ptr(nbr),
                                         _ptr(ind), _ptr(outd), _ptr(w0), _ptr(b0), _ptr(w1),
                                         _ptr(b1), float(gin_eps), fi, hid, fo, math))
    return y


def sage_conv(x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table,
              out_degree_table, neighbor_lin_weight, neighbor_lin_bias, self_lin_weight,
              math: int = FAST):
    """lib:2211-2341"""
    wl, bl, wr = map(_f32, (neighbor_lin_weight, neighbor_lin_bias, self_lin_weight))
    fo, fi = int(wl.shape[0]), int(wl.shape[1])
    x, y, n, e, coo, off, nbr, ind, outd = _conv_prologue(
        x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table, out_degree_table, fo)
    _lib.check(_lib.load().gnnb_sage_conv(n, e, _ptr(x), _ptr(y), _ptr(coo), _ptr(off), _ptr(nbr),
                                          _ptr(ind), _ptr(outd), _ptr(wl), _ptr(bl), _ptr(wr), fi,
                                          fo, math))
    return y


def pna_conv(x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table,
             out_degree_table, transform_lin_weight, transform_lin_bias, apply_lin_weight,
             apply_lin_bias, final_lin_weight, final_lin_bias, pna_avg_degree_log: float):
    """lib:1891-2157"""
    ws = list(map(_f32, (transform_lin_weight, transform_lin_bias, apply_lin_weight,
                         apply_lin_bias, final_lin_weight, final_lin_bias)))
    fi, fo = int(ws[0].shape[0]), int(ws[2].shape[0])
    x, y, n, e, coo, off, nbr, ind, outd = _conv_prologue(
        x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table, out_degree_table, fo)
    _lib.check(_lib.load().gnnb_pna_conv(n, e, _ptr(x), _ptr(y), _ptr(coo), _ptr(off), _ptr(nbr),
                                         _ptr(ind), _ptr(outd), *[_ptr(w) for w in ws],
                                         float(pna_avg_degree_log), fi, fo))
    return y


def gine_conv(x, edge_feature_table, edge_list, neighbor_table_offsets, neighbor_table,
              edge_index_table, in_degree_table, out_degree_table, edge_proj_weight,
              edge_proj_bias, mlp_0_weight, mlp_0_bias, mlp_1_weight, mlp_1_bias, gin_eps: float,
              math: int = FAST):
    """lib:1627-1742"""
    we, be, w0, b0, w1, b1 = map(_f32, (edge_proj_weight, edge_proj_bias, mlp_0_weight, mlp_0_bias,
                                        mlp_1_weight, mlp_1_bias))
    hid, fi, fo, fe = int(w0.shape[0]), int(w0.shape[1]), int(w1.shape[0]), int(we.shape[1])
    x, y, n, e, coo, off, nbr, ind, outd = _conv_prologue(
        x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table, out_degree_table, fo)
    ef, eidx = _f32(edge_feature_table), _i32(edge_index_table)
    _lib.check(_lib.load().gnnb_gine_conv(n, e, _ptr(x), _ptr(y), _ptr(ef), _ptr(coo), _ptr(off),
                                          _ptr(nbr), _ptr(eidx), _ptr(ind), _ptr(outd), _ptr(we),
                                          _ptr(be), _ptr(w0), _ptr(b0), _ptr(w1), _ptr(b1),
                                          float(gin_eps), fi, hid, fo, fe, math))
    return y


def _agg_only(name, x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table,
              out_degree_table, math):
    x = _f32(x)
    f = int(x.shape[1])
    x, y, n, e, coo, off, nbr, ind, outd = _conv_prologue(
        x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table, out_degree_table, f)
    _lib.check(getattr(_lib.load(), name)(n, e, _ptr(x), _ptr(y), _ptr(coo), _ptr(off), _ptr(nbr),
                                          _ptr(ind), _ptr(outd), f, math))
    return y


def lg_conv(x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table,
            out_degree_table, math: int = FAST):
    """lib:2398-2499"""
    return _agg_only("gnnb_lg_conv", x, edge_list, neighbor_table_offsets, neighbor_table,
                     in_degree_table, out_degree_table, math)


def simple_conv(x, edge_list, neighbor_table_offsets, neighbor_table, in_degree_table,
                out_degree_table, math: int = FAST):
    """lib:2564-2634"""
    return _agg_only("gnnb_simple_conv", x, edge_list, neighbor_table_offsets, neighbor_table,
                     in_degree_table, out_degree_table, math)


def _pool(kind: str, x):
    x = _f32(x)
    n, f = int(x.shape[0]), int(x.shape[1])
    out = _empty_like_space(x, (f,), np.float32)
    _lib.check(getattr(_lib.load(), f"gnnb_global_{kind}_pool")(n, 0, _ptr(x), _ptr(out), f))
    return out


def global_add_pool(x):
    """lib:2709-2739"""
    return _pool("add", x)


def global_mean_pool(x):
    """lib:2741-2771"""
    return _pool("mean", x)


def global_max_pool(x):
    """lib:2773-2803"""
    return _pool("max", x)
