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
    _lib.check(_lib.load().gnnb_debug_tc_gemm(C.c_void_p(A.ctypes.data), C.c_void_p(W.ctypes.data),
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
