// Stand-alone check of the tcgen05 building blocks (tc.cuh): one CTA computes
//   C[128][N] = A[128][K] . W[N][K]^T      (3xTF32, fp32 accumulate in tensor memory)
// with exactly the operand layouts, descriptors, bulk copies and barriers the fused kernel uses.
// Exposed as gnnb_debug_tc_gemm so that a parity test can pin the primitives in isolation.
// Built into libgnnb_b200_debug.so (test/diagnostic library), NOT into the product library.
#include <vector>

#include "../../include/gnnb_b200_debug.h"
#include "kernels.h"
#include "tc.cuh"

namespace gnnb {

void build_weight_image(const float *W, int N, int n_valid, int K, int ld, int col0,
                        std::vector<float> &img);   // fused_tc.cu (libgnnb_b200.so)

namespace {

constexpr int TM = 128;

__global__ void __launch_bounds__(256, 1) tc_gemm_test_kernel(const float *__restrict__ A,
                                                              const float *__restrict__ Bimg,
                                                              float *__restrict__ C, int K, int N)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KA = (K + tc::ATOM_K - 1) / tc::ATOM_K;
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *a_hi = base;
    unsigned char *a_lo = a_hi + (size_t)KA * TM * tc::ROW_BYTES;
    unsigned char *b_st = a_lo + (size_t)KA * TM * tc::ROW_BYTES;
    const uint32_t stage_bytes = 2u * (uint32_t)N * tc::ROW_BYTES;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, 128);
    if (tid == 0) {
        tc::mbar_init(&bar_full[0], 1); tc::mbar_init(&bar_full[1], 1);
        tc::mbar_init(&bar_empty[0], 1); tc::mbar_init(&bar_empty[1], 1);
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    // A -> hi / lo operand buffers (zero padded to whole K atoms)
    for (int idx = tid; idx < TM * KA * 8; idx += blockDim.x) {
        const int row = idx / (KA * 8), c4 = (idx % (KA * 8)) * 4;
        float v[4], h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j] = (c4 + j < K) ? A[(size_t)row * K + c4 + j] : 0.0f;
            h[j] = tc::tf32_hi(v[j]);
            l[j] = v[j] - h[j];
        }
        const uint32_t off = tc::canon_chunk_offset(row, c4, TM);
        *reinterpret_cast<float4 *>(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(a_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(TM, N);
        const unsigned char *bsrc = reinterpret_cast<const unsigned char *>(Bimg);
        for (int s = 0; s < 2 && s < KA; s++) {
            tc::mbar_expect_tx(&bar_full[s], stage_bytes);
            tc::bulk_g2s(b_st + (size_t)s * stage_bytes, bsrc + (size_t)s * stage_bytes, stage_bytes,
                         &bar_full[s]);
        }
        uint32_t full_par[2] = {0, 0}, empty_par[2] = {0, 0};
        for (int ka = 0; ka < KA; ka++) {
            const int s = ka & 1;
            tc::mbar_wait(&bar_full[s], full_par[s]);
            full_par[s] ^= 1;
            tc::tc_fence_after();
            const uint32_t ah = tc::smem_u32(a_hi) + (uint32_t)ka * TM * tc::ROW_BYTES;
            const uint32_t al = tc::smem_u32(a_lo) + (uint32_t)ka * TM * tc::ROW_BYTES;
            const uint32_t bh = tc::smem_u32(b_st) + (uint32_t)s * stage_bytes;
            const uint32_t bl = bh + (uint32_t)N * tc::ROW_BYTES;
#pragma unroll
            for (int k8 = 0; k8 < tc::ATOM_K / tc::MMA_K; k8++) {
                const uint32_t ko = (uint32_t)k8 * tc::MMA_K * 4;  // bytes along K inside the atom
                const uint32_t first = (ka == 0 && k8 == 0) ? 0u : 1u;
                tc::mma_tf32(tmem_d, tc::make_desc(ah + ko), tc::make_desc(bh + ko), idesc, first);
                tc::mma_tf32(tmem_d, tc::make_desc(al + ko), tc::make_desc(bh + ko), idesc, 1u);
                tc::mma_tf32(tmem_d, tc::make_desc(ah + ko), tc::make_desc(bl + ko), idesc, 1u);
            }
            tc::mma_commit(&bar_empty[s]);
            if (ka >= 1 && ka + 1 < KA) {  // refill the stage atom ka-1 used with atom ka+1
                const int sp = (ka - 1) & 1;
                tc::mbar_wait(&bar_empty[sp], empty_par[sp]);
                empty_par[sp] ^= 1;
                tc::mbar_expect_tx(&bar_full[sp], stage_bytes);
                tc::bulk_g2s(b_st + (size_t)sp * stage_bytes, bsrc + (size_t)(ka + 1) * stage_bytes,
                             stage_bytes, &bar_full[sp]);
            }
        }
        tc::mma_commit(&bar_done);
    }
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after();
    {
        const int row = 32 * (warp & 3) + lane;
        for (int c0 = (warp >> 2) * 32; c0 < N; c0 += 64) {
            float v[32];
            tc::tmem_ld32(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (c0 + j < N) C[(size_t)row * N + c0 + j] = v[j];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 128);
}


// Second diagnostic: the primitives of the tensor-core aggregation path on one tile,
//   C[128][N] = (ADJ[128][128] . X[128][F]) . W[N][F]^T
// ADJ.X as kind::f16 MMAs (ADJ bf16 K-major, X as three bf16 planes MN-major), the product moved
// to tensor memory as the hi/lo TF32 A operand with tcgen05.st, then the 3xTF32 transform with
// A from tensor memory -- exactly the operand layouts and descriptors fused_tc.cu uses.
__global__ void __launch_bounds__(256, 1) tc_agg_test_kernel(const float *__restrict__ Adj,
                                                             const float *__restrict__ X,
                                                             const float *__restrict__ Bimg,
                                                             float *__restrict__ C,
                                                             float *__restrict__ AggOut, int F, int N)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full, bar_empty, bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KA = (F + tc::ATOM_K - 1) / tc::ATOM_K, kp = KA * tc::ATOM_K;
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *ADJ = base;
    unsigned char *XP = ADJ + tc::PLANE_BYTES;            // three planes
    unsigned char *WS = XP + 3 * tc::PLANE_BYTES;         // one weight slot (N x 128 B)
    const uint32_t slot_bytes = (uint32_t)N * tc::ROW_BYTES;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        tc::mbar_init(&bar_full, 1); tc::mbar_init(&bar_empty, 1); tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    // ADJ: 8 sources per thread-iteration -> one 16-byte chunk of bf16
    for (int idx = tid; idx < TM * 16; idx += blockDim.x) {
        const int d = idx >> 4, s0 = (idx & 15) * 8;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t a = __float_as_uint(Adj[(size_t)d * TM + s0 + 2 * j]);
            const uint32_t b = __float_as_uint(Adj[(size_t)d * TM + s0 + 2 * j + 1]);
            w[j] = __byte_perm(a, b, 0x7632);
        }
        *reinterpret_cast<uint4 *>(ADJ + tc::adj_chunk_offset(d, s0)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // X planes: thread (row, half) handles every second chunk of 8 features
    {
        const int r = tid & 127;
        for (int c = tid >> 7; c < kp / 8; c += 2) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (c * 8 + j < F) ? X[(size_t)r * F + c * 8 + j] : 0.0f;
            uint4 h, m, l;
            tc::split3_pack8(v, h, m, l);
            const uint32_t off = tc::plane_chunk_offset(r, c * 8);
            *reinterpret_cast<uint4 *>(XP + off) = h;
            *reinterpret_cast<uint4 *>(XP + tc::PLANE_BYTES + off) = m;
            *reinterpret_cast<uint4 *>(XP + 2 * tc::PLANE_BYTES + off) = l;
        }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot, tmem_ahi = tmem_slot + 128, tmem_alo = tmem_slot + 256;

    if (tid == 0) {   // AGG = ADJ . (Xh + Xm + Xl): K = 128 source nodes, 16 per MMA
        const uint32_t idesc = tc::make_idesc_bf16(TM, kp, 1);
        for (int pl = 0; pl < 3; pl++)
            for (int ks = 0; ks < 8; ks++) {
                const uint32_t a = tc::smem_u32(ADJ) + (uint32_t)(ks >> 2) * tc::PLANE_BLOCK_BYTES +
                                   (uint32_t)(ks & 3) * 32u;
                const uint32_t b = tc::smem_u32(XP) + (uint32_t)pl * tc::PLANE_BYTES + (uint32_t)ks * 2048u;
                tc::mma_bf16(tmem_d, tc::make_desc(a),
                             tc::make_desc_mn(b, tc::PLANE_BLOCK_BYTES, 1024u), idesc,
                             (pl == 0 && ks == 0) ? 0u : 1u);
            }
        tc::mma_commit(&bar_done);
    }
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after();
    {   // accumulator -> (hi, lo) A operand in tensor memory, thread per row
        const int row = 32 * (warp & 3) + lane;
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        for (int c0 = (warp >> 2) * 32; c0 < kp; c0 += 64) {
            float v[32], h[32], l[32];
            tc::tmem_ld32(tmem_d + lane_base + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                h[j] = tc::tf32_hi(v[j]);
                l[j] = v[j] - h[j];
                if (AggOut != nullptr && c0 + j < F) AggOut[(size_t)row * F + c0 + j] = v[j];
            }
            tc::tmem_st32(tmem_ahi + lane_base + (uint32_t)c0, h);
            tc::tmem_st32(tmem_alo + lane_base + (uint32_t)c0, l);
        }
        tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) {   // C = A . W^T with A from tensor memory, weight units through one slot
        const uint32_t idesc = tc::make_idesc_tf32(TM, N);
        const unsigned char *bsrc = reinterpret_cast<const unsigned char *>(Bimg);
        for (int u = 0; u < 2 * KA; u++) {
            const int part = u >= KA ? 1 : 0, ka = part ? u - KA : u;
            if (u > 0) tc::mbar_wait(&bar_empty, (u - 1) & 1);
            tc::mbar_expect_tx(&bar_full, slot_bytes);
            tc::bulk_g2s(WS, bsrc + ((size_t)ka * 2 + part) * slot_bytes, slot_bytes, &bar_full);
            tc::mbar_wait(&bar_full, u & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int k8 = 0; k8 < tc::ATOM_K / tc::MMA_K; k8++) {
                const uint32_t col = (uint32_t)(ka * tc::ATOM_K + k8 * tc::MMA_K);
                const uint64_t b = tc::make_desc(tc::smem_u32(WS) + (uint32_t)k8 * tc::MMA_K * 4);
                if (part == 0) {
                    tc::mma_tf32_ts(tmem_d, tmem_ahi + col, b, idesc, (u == 0 && k8 == 0) ? 0u : 1u);
                    tc::mma_tf32_ts(tmem_d, tmem_alo + col, b, idesc, 1u);
                } else {
                    tc::mma_tf32_ts(tmem_d, tmem_ahi + col, b, idesc, 1u);
                }
            }
            tc::mma_commit(&bar_empty);
        }
        tc::mma_commit(&bar_done);
    }
    tc::mbar_wait(&bar_done, 1);
    tc::tc_fence_after();
    {
        const int row = 32 * (warp & 3) + lane;
        for (int c0 = (warp >> 2) * 32; c0 < N; c0 += 64) {
            float v[32];
            tc::tmem_ld32(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (c0 + j < N) C[(size_t)row * N + c0 + j] = v[j];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 512);
}


// Third diagnostic: issue-to-completion cycles of back-to-back tcgen05.mma instructions of the
// flavours the fused kernel uses (M = 128, N columns): 0 tf32 SS, 1 tf32 TS (A in tensor memory),
// 2 bf16 SS K-major, 3 bf16 SS with MN-major B.  One CTA, zeroed operands.
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(int flavour_in, int N, int reps, int nacc,
                                                             long long *cycles)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    int flavour = flavour_in;
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = tid; i < 3 * tc::PLANE_BYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) { tc::mbar_init(&bar_done, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    {   // zero the A region of tensor memory (columns 128..383)
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; j++) z[j] = 0.0f;
        for (int c = 128; c < 384; c += 32) tc::tmem_st32(tmem_d + ((uint32_t)(32 * warp) << 16) + c, z);
        tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const bool uniform = flavour >= 10;
    if (uniform) flavour -= 10;
    // broadcast from lane 0 so that the compiler treats the values as warp-uniform (uniform
    // registers, no per-instruction R2UR)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
    const uint32_t a_u = __shfl_sync(0xffffffffu, tc::smem_u32(base), 0);
    if (uniform ? warp_u == 0 : tid == 0) {
        const uint32_t tmem_d = tmem_u;
        const uint32_t a = a_u, b = a + tc::PLANE_BYTES;
        const uint32_t id_tf = tc::make_idesc_tf32(TM, N);
        const uint32_t id_bk = tc::make_idesc_bf16(TM, N, 0), id_bm = tc::make_idesc_bf16(TM, N, 1);
        const long long t0 = clock64();
        if (!uniform) {
            for (int r = 0; r < reps; r++) {
                const uint32_t ko = (uint32_t)(r & 3) * 32u;
                const uint32_t tmem_d = tmem_u + (uint32_t)(r % nacc) * (uint32_t)N;
                if (flavour == 0)
                    tc::mma_tf32(tmem_d, tc::make_desc(a + ko), tc::make_desc(b + ko), id_tf, r ? 1u : 0u);
                else if (flavour == 1)
                    tc::mma_tf32_ts(tmem_d, tmem_u + 256 + (uint32_t)(r & 15) * 8u, tc::make_desc(b + ko), id_tf,
                                    r ? 1u : 0u);
                else if (flavour == 2)
                    tc::mma_bf16(tmem_d, tc::make_desc(a + ko), tc::make_desc(b + ko), id_bk, r ? 1u : 0u);
                else
                    tc::mma_bf16(tmem_d, tc::make_desc(a + ko),
                                 tc::make_desc_mn(b + (uint32_t)(r & 7) * 2048u, tc::PLANE_BLOCK_BYTES, 1024u),
                                 id_bm, r ? 1u : 0u);
            }
        } else {   // the whole warp runs the loop, one elected lane issues
            for (int r0 = 0; r0 < reps; r0 += 4) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int r = r0 + k;
                    const uint32_t ko = (uint32_t)k * 32u;
                    if (flavour == 0) {
                        const uint64_t da = tc::make_desc(a + ko), db = tc::make_desc(b + ko);
                        if (tc::elect_one()) tc::mma_tf32(tmem_d, da, db, id_tf, r ? 1u : 0u);
                    } else if (flavour == 1) {
                        const uint64_t db = tc::make_desc(b + ko);
                        const uint32_t ta = tmem_d + 128 + (uint32_t)((r0 & 12) + k) * 8u;
                        if (tc::elect_one()) tc::mma_tf32_ts(tmem_d, ta, db, id_tf, r ? 1u : 0u);
                    } else if (flavour == 2) {
                        const uint64_t da = tc::make_desc(a + ko), db = tc::make_desc(b + ko);
                        if (tc::elect_one()) tc::mma_bf16(tmem_d, da, db, id_bk, r ? 1u : 0u);
                    } else {
                        const uint64_t da = tc::make_desc(a + ko);
                        const uint64_t db = tc::make_desc_mn(b + (uint32_t)((r0 & 4) + k) * 2048u,
                                                             tc::PLANE_BLOCK_BYTES, 1024u);
                        if (tc::elect_one()) tc::mma_bf16(tmem_d, da, db, id_bm, r ? 1u : 0u);
                    }
                }
            }
        }
        const long long t1 = clock64();
        if (!uniform) tc::mma_commit(&bar_done);          // (elect.sync needs the whole warp)
        else if (tc::elect_one()) tc::mma_commit(&bar_done);
        tc::mbar_wait(&bar_done, 0);
        const long long t2 = clock64();
        if (tid == 0) {
            cycles[0] = t1 - t0;
            cycles[1] = t2 - t0;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 512);
}


// Lean issue loop: the whole of warp 0 runs warp-uniform code (values broadcast from lane 0 so
// that the compiler keeps them in uniform registers), 16 MMAs unrolled per iteration with
// compile-time offsets, one elected lane executes the instruction.
template <int FLAVOUR>
__global__ void __launch_bounds__(128, 1) tc_mma_rate_lean_kernel(int M, int N, int reps, long long *cycles)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = tid; i < 3 * tc::PLANE_BYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) { tc::mbar_init(&bar_done, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_slot, 0);
    {
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; j++) z[j] = 0.0f;
        for (int c = 128; c < 384; c += 32) tc::tmem_st32(tmem_d + ((uint32_t)(32 * warp) << 16) + c, z);
        tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 0) {
        const uint32_t a = __shfl_sync(0xffffffffu, tc::smem_u32(base), 0), b = a + tc::PLANE_BYTES;
        const uint32_t idesc = FLAVOUR <= 1 ? tc::make_idesc_tf32(M, N) : tc::make_idesc_bf16(M, N, FLAVOUR == 3);
        const uint64_t da0 = tc::make_desc(a), db0 = tc::make_desc(b);
        const uint64_t dm0 = tc::make_desc_mn(b, tc::PLANE_BLOCK_BYTES, 1024u);
        const bool leader = tc::elect_one();
        const long long t0 = clock64();
        for (int r0 = 0; r0 < reps; r0 += 16) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t acc = (r0 + k) ? 1u : 0u;
                if (FLAVOUR == 0) {
                    if (leader) tc::mma_tf32(tmem_d, da0 + (uint64_t)((k & 3) * 2), db0 + (uint64_t)((k & 3) * 2), idesc, acc);
                } else if (FLAVOUR == 1) {
                    if (leader) tc::mma_tf32_ts(tmem_d, tmem_d + 128 + k * 8, db0 + (uint64_t)((k & 3) * 2), idesc, acc);
                } else if (FLAVOUR == 2) {
                    if (leader) tc::mma_bf16(tmem_d, da0 + (uint64_t)((k & 3) * 2), db0 + (uint64_t)((k & 3) * 2), idesc, acc);
                } else {
                    if (leader) tc::mma_bf16(tmem_d, da0 + (uint64_t)((k & 3) * 2), dm0 + (uint64_t)((k & 7) * 128), idesc, acc);
                }
            }
        }
        const long long t1 = clock64();
        if (leader) tc::mma_commit(&bar_done);
        tc::mbar_wait(&bar_done, 0);
        const long long t2 = clock64();
        if (tid == 0) {
            cycles[0] = t1 - t0;
            cycles[1] = t2 - t0;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 512);
}


// Probe for the next kernel generation: a kind::f16 MMA whose A operand (bf16) lives in TENSOR
// MEMORY.  C[128][N] = bf16(A)[128][K] . bf16(B)[N][K]^T, fp32 accumulate.  `variant` selects the
// hypothesis for how 16-bit A elements sit in the 32-bit TMEM cells of a lane:
//   0  packed pairs: cell j = { k = 2j in the low half, k = 2j + 1 in the high half }, 8 cells per MMA
//   1  one element per cell, in the low 16 bits, 16 cells per MMA
//   2  one element per cell, in the high 16 bits (an fp32 value truncated in place), 16 cells per MMA
// tools/tmem_bf16_probe.py prints which one reproduces the fp64 product of the truncated operands.
__global__ void __launch_bounds__(128, 1) tc_bf16_ts_test_kernel(const float *__restrict__ A,
                                                                 const float *__restrict__ B,
                                                                 float *__restrict__ C, int K, int N,
                                                                 int variant)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *b_s = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KB = (K + 63) / 64;   // 64-element K atoms of the bf16 B operand
    constexpr uint32_t TM_A = 128;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) {
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    // B -> bf16 (truncation), K-major SWIZZLE_128B atoms of 64 elements, rows at a 128-byte pitch
    for (int idx = tid; idx < TM * KB * 8; idx += blockDim.x) {
        const int n = idx / (KB * 8), c8 = (idx % (KB * 8)) * 8;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float lo = (n < N && c8 + 2 * j < K) ? B[(size_t)n * K + c8 + 2 * j] : 0.0f;
            const float hi = (n < N && c8 + 2 * j + 1 < K) ? B[(size_t)n * K + c8 + 2 * j + 1] : 0.0f;
            w[j] = __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
        }
        *reinterpret_cast<uint4 *>(b_s + tc::adj_chunk_offset(n, c8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    {   // this thread's row of A -> tensor memory (K <= 128, K % 32 == 0)
        const float *arow = A + (size_t)(32 * (warp & 3) + lane) * K;
        const int cells = variant == 0 ? K / 2 : K;
        for (int c0 = 0; c0 < cells; c0 += 16) {
            float cell[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t u;
                if (variant == 0)
                    u = __byte_perm(__float_as_uint(arow[2 * (c0 + j)]), __float_as_uint(arow[2 * (c0 + j) + 1]),
                                    0x7632);
                else if (variant == 1)
                    u = __float_as_uint(arow[c0 + j]) >> 16;
                else
                    u = __float_as_uint(arow[c0 + j]) & 0xffff0000u;
                cell[j] = __uint_as_float(u);
            }
            tc::tmem_st16(tmem_base + TM_A + lane_base + (uint32_t)c0, cell);
        }
        tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 0) {
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::make_idesc_bf16(TM, N, 0);
        const uint32_t cells_per_mma = variant == 0 ? 8u : 16u;
        const uint32_t b_addr = tc::smem_u32(b_s);
        for (int ks = 0; ks < K / 16; ks++) {   // 16 elements of K per MMA; 4 k-steps per 64-element atom
            const uint64_t bd = tc::make_desc(b_addr + (uint32_t)(ks >> 2) * tc::PLANE_BLOCK_BYTES) +
                                (uint64_t)(2 * (ks & 3));
            if (leader)
                tc::mma_bf16_ts(tmem_base, tmem_base + TM_A + (uint32_t)ks * cells_per_mma, bd, idesc,
                                ks == 0 ? 0u : 1u);
        }
        if (leader) tc::mma_commit(&bar_done);
    }
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after();
    {
        const int row = 32 * (warp & 3) + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tc::tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (c0 + j < N) C[(size_t)row * N + c0 + j] = v[j];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

}  // namespace
}  // namespace gnnb

using namespace gnnb;

// Diagnostic entry point: C[128][N] = A[128][K] . W[N][K]^T on the tensor cores (host buffers).
extern "C" int gnnb_debug_tc_gemm(const float *A, const float *W, float *C, int K, int N)
{
    GNNB_REQUIRE(A && W && C, "null argument");
    GNNB_REQUIRE(K >= 1 && K <= 128 && N >= 16 && N <= 128 && N % 16 == 0,
                 "tc gemm: 1 <= K <= 128, N a multiple of 16 up to 128");
    const int KA = (K + tc::ATOM_K - 1) / tc::ATOM_K;
    std::vector<float> img;
    build_weight_image(W, N, N, K, K, 0, img);
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    GNNB_CUDA(cudaMalloc(&dA, sizeof(float) * 128 * K));
    GNNB_CUDA(cudaMalloc(&dB, sizeof(float) * img.size()));
    GNNB_CUDA(cudaMalloc(&dC, sizeof(float) * 128 * N));
    GNNB_CUDA(cudaMemcpy(dA, A, sizeof(float) * 128 * K, cudaMemcpyHostToDevice));
    GNNB_CUDA(cudaMemcpy(dB, img.data(), sizeof(float) * img.size(), cudaMemcpyHostToDevice));
    const size_t smem = 1024 + (size_t)2 * KA * TM * tc::ROW_BYTES + (size_t)2 * 2 * N * tc::ROW_BYTES;
    GNNB_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    tc_gemm_test_kernel<<<1, 256, smem>>>(dA, dB, dC, K, N);
    GNNB_CUDA(cudaGetLastError());
    GNNB_CUDA(cudaDeviceSynchronize());
    GNNB_CUDA(cudaMemcpy(C, dC, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return GNNB_OK;
}

// Diagnostic entry point: C[128][N] = (Adj[128][128] . X[128][F]) . W[N][F]^T, Adj holding small
// non-negative integers (edge multiplicities); agg (optional, [128][F]) receives Adj . X.
extern "C" int gnnb_debug_tc_agg_gemm(const float *Adj, const float *X, const float *W, float *C,
                                      float *agg, int F, int N)
{
    GNNB_REQUIRE(Adj && X && W && C, "null argument");
    GNNB_REQUIRE(F >= 1 && F <= 128 && N >= 16 && N <= 128 && N % 16 == 0,
                 "tc agg gemm: 1 <= F <= 128, N a multiple of 16 up to 128");
    std::vector<float> img;
    build_weight_image(W, N, N, F, F, 0, img);
    float *dAdj = nullptr, *dX = nullptr, *dB = nullptr, *dC = nullptr, *dAgg = nullptr;
    GNNB_CUDA(cudaMalloc(&dAdj, sizeof(float) * 128 * 128));
    GNNB_CUDA(cudaMalloc(&dX, sizeof(float) * 128 * F));
    GNNB_CUDA(cudaMalloc(&dB, sizeof(float) * img.size()));
    GNNB_CUDA(cudaMalloc(&dC, sizeof(float) * 128 * N));
    GNNB_CUDA(cudaMalloc(&dAgg, sizeof(float) * 128 * F));
    GNNB_CUDA(cudaMemcpy(dAdj, Adj, sizeof(float) * 128 * 128, cudaMemcpyHostToDevice));
    GNNB_CUDA(cudaMemcpy(dX, X, sizeof(float) * 128 * F, cudaMemcpyHostToDevice));
    GNNB_CUDA(cudaMemcpy(dB, img.data(), sizeof(float) * img.size(), cudaMemcpyHostToDevice));
    const size_t smem = 1024 + (size_t)4 * tc::PLANE_BYTES + (size_t)N * tc::ROW_BYTES;
    GNNB_CUDA(cudaFuncSetAttribute(tc_agg_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    tc_agg_test_kernel<<<1, 256, smem>>>(dAdj, dX, dB, dC, agg ? dAgg : nullptr, F, N);
    GNNB_CUDA(cudaGetLastError());
    GNNB_CUDA(cudaDeviceSynchronize());
    GNNB_CUDA(cudaMemcpy(C, dC, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
    if (agg) GNNB_CUDA(cudaMemcpy(agg, dAgg, sizeof(float) * 128 * F, cudaMemcpyDeviceToHost));
    cudaFree(dAdj); cudaFree(dX); cudaFree(dB); cudaFree(dC); cudaFree(dAgg);
    return GNNB_OK;
}

// Diagnostic: cycles[0] = cycles to ISSUE `reps` back-to-back MMAs, cycles[1] = until they completed.
extern "C" int gnnb_debug_tc_mma_rate(int flavour, int N, int reps, long long *cycles)
{
    const int M = flavour >= 1000 ? 64 : 128;   // thousands digit: 64-row MMAs (lean kernels only)
    flavour %= 1000;
    const int nacc = flavour / 100 > 0 ? flavour / 100 : 1;   // hundreds digit: accumulators to alternate
    flavour %= 100;
    GNNB_REQUIRE(M == 128 || flavour >= 20, "64-row MMAs: lean flavours (20..23) only");
    GNNB_REQUIRE(cycles != nullptr && flavour >= 0 && flavour <= 23 && N >= 16 && N <= 256 && N % 16 == 0 &&
                 reps >= 1, "bad argument");
    long long *d = nullptr;
    GNNB_CUDA(cudaMalloc(&d, 2 * sizeof(long long)));
    const size_t smem = 1024 + (size_t)3 * tc::PLANE_BYTES;
    GNNB_CUDA(cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    if (flavour >= 20) {
        switch (flavour - 20) {
        case 0:
            GNNB_CUDA(cudaFuncSetAttribute(tc_mma_rate_lean_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tc_mma_rate_lean_kernel<0><<<1, 128, smem>>>(M, N, reps, d); break;
        case 1:
            GNNB_CUDA(cudaFuncSetAttribute(tc_mma_rate_lean_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tc_mma_rate_lean_kernel<1><<<1, 128, smem>>>(M, N, reps, d); break;
        case 2:
            GNNB_CUDA(cudaFuncSetAttribute(tc_mma_rate_lean_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tc_mma_rate_lean_kernel<2><<<1, 128, smem>>>(M, N, reps, d); break;
        default:
            GNNB_CUDA(cudaFuncSetAttribute(tc_mma_rate_lean_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tc_mma_rate_lean_kernel<3><<<1, 128, smem>>>(M, N, reps, d); break;
        }
    } else
        tc_mma_rate_kernel<<<1, 128, smem>>>(flavour, N, reps, nacc, d);
    GNNB_CUDA(cudaGetLastError());
    GNNB_CUDA(cudaDeviceSynchronize());
    GNNB_CUDA(cudaMemcpy(cycles, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return GNNB_OK;
}

// Probe: kind::f16 MMA with a bf16 A operand in tensor memory (see tc_bf16_ts_test_kernel).
extern "C" int gnnb_debug_tc_bf16_ts(const float *A, const float *B, float *C, int K, int N, int variant)
{
    GNNB_REQUIRE(A && B && C, "null argument");
    GNNB_REQUIRE(K >= 32 && K <= 128 && K % 32 == 0 && N >= 16 && N <= 128 && N % 16 == 0,
                 "tc bf16 ts probe: K a multiple of 32 up to 128, N a multiple of 16 up to 128");
    GNNB_REQUIRE(variant >= 0 && variant <= 2, "tc bf16 ts probe: variant 0, 1 or 2");
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
    GNNB_CUDA(cudaMalloc(&dA, sizeof(float) * 128 * K));
    GNNB_CUDA(cudaMalloc(&dB, sizeof(float) * N * K));
    GNNB_CUDA(cudaMalloc(&dC, sizeof(float) * 128 * N));
    GNNB_CUDA(cudaMemcpy(dA, A, sizeof(float) * 128 * K, cudaMemcpyHostToDevice));
    GNNB_CUDA(cudaMemcpy(dB, B, sizeof(float) * N * K, cudaMemcpyHostToDevice));
    const size_t smem = 1024 + (size_t)2 * tc::PLANE_BLOCK_BYTES;
    GNNB_CUDA(cudaFuncSetAttribute(tc_bf16_ts_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    tc_bf16_ts_test_kernel<<<1, 128, smem>>>(dA, dB, dC, K, N, variant);
    GNNB_CUDA(cudaGetLastError());
    GNNB_CUDA(cudaDeviceSynchronize());
    GNNB_CUDA(cudaMemcpy(C, dC, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return GNNB_OK;
}
