"""tcgen05 building blocks in isolation: the one-CTA 3xTF32 tensor-core GEMM
(gnnb_debug_tc_gemm) against float64 numpy.  fp32-grade accuracy is the point of the
error-compensated split, so the bound is tight (2e-6 relative to the row/column norms)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def tc_gemm(A, W):
    from gnn_builder_b200 import _lib

    A = np.ascontiguousarray(A, np.float32)
    W = np.ascontiguousarray(W, np.float32)
    K, N = A.shape[1], W.shape[0]
    out = np.empty((128, N), np.float32)
    _lib.check(_lib.load_debug().gnnb_debug_tc_gemm(C.c_void_p(A.ctypes.data), C.c_void_p(W.ctypes.data),
                                              C.c_void_p(out.ctypes.data), K, N))
    return out


@pytest.mark.parametrize("K,N", [(128, 128), (32, 64), (11, 128), (64, 64), (128, 16), (100, 48),
                                 (8, 32), (96, 112)])
def test_tc_gemm_matches_fp64(K, N):
    rng = np.random.default_rng(K * 1000 + N)
    A = rng.uniform(-1, 1, (128, K)).astype(np.float32)
    W = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    got = tc_gemm(A, W)
    scale = np.abs(A).astype(np.float64) @ np.abs(W).astype(np.float64).T
    assert np.max(np.abs(got - ref) / scale) < 2e-6


def test_tc_gemm_integer_data_exact():
    # the reference's linear test uses integer-valued data and exact comparison (test.cpp:678-745)
    A = (np.arange(128)[:, None] % 7 + np.arange(64)[None, :] % 5).astype(np.float32)
    W = (np.arange(48)[:, None] % 3 - np.arange(64)[None, :] % 4).astype(np.float32)
    assert np.array_equal(tc_gemm(A, W), A @ W.T)


def test_tc_gemm_wide_dynamic_range():
    rng = np.random.default_rng(3)
    A = (rng.standard_normal((128, 128)) * 10.0 ** rng.integers(-3, 4, (128, 128))).astype(np.float32)
    W = (rng.standard_normal((128, 128)) * 10.0 ** rng.integers(-3, 4, (128, 128))).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(W).astype(np.float64).T
    assert np.max(np.abs(tc_gemm(A, W) - ref) / scale) < 5e-6


def tc_agg_gemm(Adj, X, W):
    from gnn_builder_b200 import _lib

    Adj = np.ascontiguousarray(Adj, np.float32)
    X = np.ascontiguousarray(X, np.float32)
    W = np.ascontiguousarray(W, np.float32)
    F, N = X.shape[1], W.shape[0]
    out = np.empty((128, N), np.float32)
    agg = np.empty((128, F), np.float32)
    _lib.check(_lib.load_debug().gnnb_debug_tc_agg_gemm(
        C.c_void_p(Adj.ctypes.data), C.c_void_p(X.ctypes.data), C.c_void_p(W.ctypes.data),
        C.c_void_p(out.ctypes.data), C.c_void_p(agg.ctypes.data), F, N))
    return out, agg


def _random_adjacency(rng, density=0.03, self_loops=True):
    adj = (rng.random((128, 128)) < density).astype(np.float32)
    adj += (rng.random((128, 128)) < 0.002) * 2.0          # a few multi-edges
    if self_loops:
        adj += np.eye(128, dtype=np.float32)
    return adj.astype(np.float32)


@pytest.mark.parametrize("F,N", [(128, 128), (11, 128), (32, 64), (64, 16), (9, 64), (80, 80),
                                 (100, 48)])
def test_tc_aggregation_and_tmem_operand_match_fp64(F, N):
    """ADJ.X on bf16x3 planes (MN-major B operand) followed by the 3xTF32 transform with the A
    operand in tensor memory: both stages must be fp32-grade."""
    rng = np.random.default_rng(F * 131 + N)
    adj = _random_adjacency(rng)
    X = rng.uniform(-1, 1, (128, F)).astype(np.float32)
    W = rng.uniform(-1, 1, (N, F)).astype(np.float32)
    got, agg = tc_agg_gemm(adj, X, W)
    agg_ref = adj.astype(np.float64) @ X.astype(np.float64)
    agg_scale = adj.astype(np.float64) @ np.abs(X).astype(np.float64) + 1e-30
    assert np.max(np.abs(agg - agg_ref) / agg_scale) < 1e-6
    ref = agg_ref @ W.astype(np.float64).T
    scale = np.abs(agg_ref) @ np.abs(W).astype(np.float64).T + 1e-30
    assert np.max(np.abs(got - ref) / scale) < 5e-6


def test_tc_aggregation_integer_data_exact():
    rng = np.random.default_rng(11)
    adj = _random_adjacency(rng, density=0.05)
    X = rng.integers(-8, 9, (128, 64)).astype(np.float32)
    W = rng.integers(-3, 4, (32, 64)).astype(np.float32)
    got, agg = tc_agg_gemm(adj, X, W)
    assert np.array_equal(agg, adj @ X)
    assert np.array_equal(got, (adj @ X) @ W.T)


def test_tc_aggregation_wide_dynamic_range():
    rng = np.random.default_rng(5)
    adj = _random_adjacency(rng)
    X = (rng.standard_normal((128, 128)) * 10.0 ** rng.integers(-3, 4, (128, 128))).astype(np.float32)
    W = np.eye(128, dtype=np.float32)
    got, agg = tc_agg_gemm(adj, X, W)
    ref = adj.astype(np.float64) @ X.astype(np.float64)
    scale = adj.astype(np.float64) @ np.abs(X).astype(np.float64) + 1e-30
    assert np.max(np.abs(agg - ref) / scale) < 1e-6
    assert np.max(np.abs(got - ref) / scale) < 2e-6
