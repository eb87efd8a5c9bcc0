"""Dataset statistics helpers -- mirror of ``gnnbuilder/utils.py:9-115`` without the PyG
dependency.  A *dataset* is a ``GraphBatch`` or any iterable of graph objects exposing ``x``
([n, F]) and ``edge_index`` ([2, E]) (numpy arrays or torch tensors), like PyG ``Data``."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .data import GraphBatch


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def iter_graphs(dataset):
    """Yield (x float32 [n,F], coo int32 [E,2], y or None) for every graph of a dataset."""
    if isinstance(dataset, GraphBatch):
        for g in range(dataset.n_graphs):
            x, coo = dataset.graph(g)
            yield x, coo, None
        return
    for data in dataset:
        if isinstance(data, (tuple, list)):
            x, coo = data[0], data[1]
            y = data[2] if len(data) > 2 else None
            yield np.asarray(x, np.float32), np.asarray(coo, np.int32).reshape(-1, 2), y
            continue
        x = _np(data.x).astype(np.float32)
        ei = _np(data.edge_index)
        coo = np.ascontiguousarray(ei.T.astype(np.int32)) if ei.shape[0] == 2 else ei.astype(np.int32)
        yield x, coo, getattr(data, "y", None)


def compute_max_nodes_and_edges(dataset):
    """utils.py:9-15"""
    mn = me = 0
    for x, coo, _ in iter_graphs(dataset):
        mn, me = max(mn, x.shape[0]), max(me, coo.shape[0])
    return mn, me


def compute_average_nodes_and_edges(dataset, round_val: bool = True):
    """utils.py:18-31"""
    n = e = c = 0
    for x, coo, _ in iter_graphs(dataset):
        n, e, c = n + x.shape[0], e + coo.shape[0], c + 1
    an, ae = n / c, e / c
    return (int(round(an)), int(round(ae))) if round_val else (an, ae)


def compute_median_nodes_and_edges(dataset, round_val: bool = True):
    """utils.py:34-46"""
    ns, es = [], []
    for x, coo, _ in iter_graphs(dataset):
        ns.append(x.shape[0])
        es.append(coo.shape[0])
    return int(np.median(ns)), int(np.median(es))


def compute_degree(x, coo):
    """utils.py:49-57 -> (in_degree, out_degree) lists"""
    n = x.shape[0]
    return (np.bincount(coo[:, 1], minlength=n).astype(float).tolist(),
            np.bincount(coo[:, 0], minlength=n).astype(float).tolist())


def compute_average_degree(dataset, round_val=True):
    """utils.py:60-70"""
    acc = c = 0
    for x, coo, _ in iter_graphs(dataset):
        acc += np.mean(compute_degree(x, coo)[0])
        c += 1
    acc /= c
    return int(np.ceil(acc)) if round_val else acc


def compute_median_degree(dataset):
    """utils.py:73-78"""
    meds = [np.median(compute_degree(x, coo)[0]) for x, coo, _ in iter_graphs(dataset)]
    return int(np.ceil(np.median(meds)))


def compute_in_deg_histogram(dataset):
    """utils.py:81-96 (the histogram PyG's PNAConv derives its ``delta`` from)"""
    degs = [np.bincount(coo[:, 1], minlength=x.shape[0]) for x, coo, _ in iter_graphs(dataset)]
    mx = max(int(d.max()) for d in degs)
    hist = np.zeros(mx + 1, np.int64)
    for d in degs:
        hist += np.bincount(d, minlength=mx + 1)
    return hist


def layer_param_name_combiner(layer_name, param_name):
    """utils.py:99-100"""
    return f"{layer_name}_{param_name.replace('.', '_')}"


def serialize_tensor(param, fp: Path, np_type=np.float32):
    """utils.py:113-115"""
    _np(param).astype(np_type).tofile(fp)
