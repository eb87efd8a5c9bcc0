/*
 * gnnb_b200_debug.h -- diagnostics of libgnnb_b200_debug.so (NOT part of the product ABI): one-CTA
 * probes of the tcgen05 building blocks (csrc/tc.cuh) used by tests/test_gpu_tc.py and tools/.
 * The debug library links against libgnnb_b200.so.
 */
#ifndef GNNB_B200_DEBUG_H
#define GNNB_B200_DEBUG_H

#ifdef __cplusplus
extern "C" {
#endif

/* One-CTA tensor-core GEMM C[128][N] = A[128][K] . W[N][K]^T (tcgen05, 3xTF32) built from the
 * same primitives as the fused kernel's node transform; host buffers; K <= 128, N % 16 == 0. */
int gnnb_debug_tc_gemm(const float *A, const float *W, float *C, int K, int N);
/* One-tile check of the tensor-core aggregation path: C[128][N] = (Adj . X) . W^T where Adj
 * [128][128] holds edge multiplicities (small non-negative integers), X is [128][F]; the
 * aggregation runs as bf16x3 MMAs and the transform takes its A operand from tensor memory, as in
 * the fused kernel.  agg (optional, [128][F]) receives Adj . X.  Host buffers. */
/* Cycles to issue / complete `reps` back-to-back tcgen05.mma (M = 128, N columns); flavour 0 tf32
 * with both operands in shared memory, 1 tf32 with A in tensor memory, 2 bf16 K-major, 3 bf16 with
 * an MN-major B operand; + 20 = lean warp-uniform issue loop; + 1000 = M = 64 instead of 128 (lean
 * flavours only).  cycles[2]. */
int gnnb_debug_tc_mma_rate(int flavour, int N, int reps, long long *cycles);
int gnnb_debug_tc_agg_gemm(const float *Adj, const float *X, const float *W, float *C, float *agg,
                           int F, int N);

/* Probe for a kind::f16 MMA with a bf16 A operand in tensor memory: C[128][N] = bf16(A)[128][K] .
 * bf16(B)[N][K]^T; `variant` = the assumed layout of 16-bit elements in the 32-bit TMEM cells
 * (0 packed pairs, 1 low half, 2 high half; tools/tmem_bf16_probe.py).  Host buffers. */
int gnnb_debug_tc_bf16_ts(const float *A, const float *B, float *C, int K, int N, int variant);

#ifdef __cplusplus
}
#endif
#endif /* GNNB_B200_DEBUG_H */
