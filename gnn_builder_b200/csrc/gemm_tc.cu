// Node transform of the layerwise path on the 5th-generation tensor cores:
//
//   C[M][N] = act( A1[M][K1] . W1^T (+ A2[M][K2] . W2^T) + bias (+ skip) )       (gemm.cu's contract)
//
// for M in the millions (all nodes of a batch or of one large graph), error-compensated 3xTF32
// (tc.cuh) so results stay fp32-grade.  Persistent CTAs (one per SM) walk 128-row tiles; the K
// dimension streams through in 32-wide "atoms":
//
//   8 worker warps   cp.async the raw fp32 A atom (128 rows x 128 B, XOR-swizzled 16-byte chunks) into
//                    a 6-deep shared-memory ring, five atoms ahead and across tile boundaries; for
//                    the atom at the head each thread takes half of "its" row, splits it into
//                    TF32 hi / lo parts in registers and writes them to tensor memory (tcgen05.st,
//                    2 stages), then arrives on the stage's mbarrier.  One tile later the same
//                    threads run the epilogue: accumulator row -> bias / skip / activation -> C.
//   1 producer warp  streams the weight atoms (pre-split, pre-swizzled images built once per
//                    Linear, 2 x N x 128 B per atom) L2 -> shared memory with cp.async.bulk into a
//                    4-slot ring, in the order the MMAs consume them.
//   1 issuer warp    warp-uniform code, one elected lane: per atom 8 + 4 tcgen05.mma.kind::tf32 with
//                    the A operand in tensor memory, commits to the ring / stage / accumulator
//                    mbarriers.  The accumulator is double buffered (2 x N columns) so the epilogue
//                    of tile t overlaps the main loop of tile t+1.
//
// HBM roofline: 4 M (K1 + K2 + N) bytes; the weight images stay in L2.
#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "tc.cuh"

namespace gnnb {

namespace {

constexpr int TM = 128;
constexpr int NWORK = 256;                 // worker threads
constexpr int NTHREADS = NWORK + 64;       // + producer warp + issuer warp
constexpr int ASTAGES = 6;                 // raw A atoms in flight (shared memory)
constexpr int A_STAGE_BYTES = TM * tc::ROW_BYTES;   // 16 KB
constexpr int WSLOTS = 4;
constexpr int MAX_TSTAGES = 4;             // A operand stages in tensor memory: 4 when the two
                                           // accumulators leave 256 columns free, else 2; stage s
                                           // lives at [512 - 64 (s + 1), +64): hi 32 columns, lo 32

struct TcGemmParams {
    const float *A[2];
    int lda[2], K[2], KA[2];
    const float *img[2];
    const float *bias, *skip;
    int ldskip, act;
    float *C;
    int ldc, M, N, Npad, n_tiles;
    int tstages;       // 2 or 4
    // expand mode: operand 1 holds statistics [M][K[1]] standing for [A | amp_v A | att_v A]:
    // every physical atom of operand 1 is handed to the MMAs three times (scaled by 1, amp, att)
    // against the weight atoms ka, KA[1] + ka, 2 KA[1] + ka.  rep1 = 3 then, else 1.
    int rep1;
    const int32_t *expand_deg;
    float expand_delta;
};

struct Bars {
    uint64_t w_full[WSLOTS], w_empty[WSLOTS];
    uint64_t a_full[MAX_TSTAGES], a_free[MAX_TSTAGES];
    uint64_t d_full[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_dst), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(bar)) : "memory");
}

// position of the global atom stream: tile, operand, K atom
struct AtomPos {
    int tile, op, ka;
    __device__ __forceinline__ void next(const TcGemmParams &p, int stride)
    {
        if (++ka == p.KA[op]) {
            ka = 0;
            if (op == 0 && p.KA[1] > 0) op = 1;
            else { op = 0; tile += stride; }
        }
    }
};

// one thread's share of the cp.async copies of an atom: 16-byte chunk (tid & 7) of rows
// (tid >> 3) + 32 j, destination chunk XOR-swizzled with the row so that thread-per-row reads of
// the stage are bank-conflict free.  Out-of-range rows / columns are zero filled (src_bytes = 0).
__device__ __forceinline__ void issue_atom(const TcGemmParams &p, const AtomPos &a, uint32_t stage_addr,
                                           int tid)
{
    if (a.tile < p.n_tiles) {
        const float *A = p.A[a.op];
        const int lda = p.lda[a.op], K = p.K[a.op];
        const int chunk = tid & 7, k = a.ka * tc::ATOM_K + chunk * 4;
        const bool vec = (lda & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = (tid >> 3) + 32 * j;
            const int64_t row = (int64_t)a.tile * TM + r;
            const uint32_t dst = stage_addr + (uint32_t)r * tc::ROW_BYTES + (uint32_t)((chunk ^ (r & 7)) << 4);
            if (vec) {
                const int left = K - k;   // floats available from column k
                const uint32_t bytes = (row < p.M && left > 0) ? (left >= 4 ? 16u : (uint32_t)left * 4u) : 0u;
                const float *src = A + (size_t)(row < p.M ? row : 0) * lda + (left > 0 ? k : 0);
                cp_async16(dst, src, bytes);
            } else {   // unaligned rows: plain loads + a shared-memory store
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    v[q] = (row < p.M && k + q < K) ? __ldg(A + (size_t)row * lda + k + q) : 0.0f;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "f"(v[0]), "f"(v[1]),
                             "f"(v[2]), "f"(v[3])
                             : "memory");
            }
        }
    }
    cp_async_commit();
}

__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const __grid_constant__ TcGemmParams p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t slot_stride = (uint32_t)p.Npad * tc::ROW_BYTES;
    Bars &bars = *reinterpret_cast<Bars *>(base + ASTAGES * A_STAGE_BYTES + WSLOTS * slot_stride);
    const int tid = threadIdx.x;
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t a_ring = __shfl_sync(0xffffffffu, tc::smem_u32(base), 0);
    const uint32_t w_ring = a_ring + ASTAGES * A_STAGE_BYTES;

    if (warp_u == 0) tc::tmem_alloc(&bars.tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < WSLOTS; i++) { tc::mbar_init(&bars.w_full[i], 1); tc::mbar_init(&bars.w_empty[i], 1); }
        for (int i = 0; i < MAX_TSTAGES; i++) { tc::mbar_init(&bars.a_full[i], NWORK); tc::mbar_init(&bars.a_free[i], 1); }
        tc::mbar_init(&bars.d_full[0], 1);
        tc::mbar_init(&bars.d_full[1], 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, bars.tmem_slot, 0);
    const int atoms_per_tile = p.KA[0] + p.KA[1];            // physical atoms (shared-memory ring)
    const int vatoms_per_tile = p.KA[0] + p.rep1 * p.KA[1];   // hand-offs to the MMAs
    const int stride = gridDim.x;

    if (warp_u == 8) {
        // ------------------------------------------------------------ weight producer
        const bool leader = tc::elect_one();
        uint32_t prod = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += stride)
            for (int op = 0; op < 2; op++) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(p.img[op]);
                const int rep = op == 1 ? p.rep1 : 1;
                for (int pa = 0; pa < p.KA[op]; pa++)
                    for (int r = 0; r < rep; r++)
                        for (int part = 0; part < 2; part++) {   // weight atom r KA + pa: hi, then lo
                            const size_t u = (size_t)(r * p.KA[op] + pa) * 2 + part;
                            const uint32_t s = prod & (WSLOTS - 1), use = prod / WSLOTS;
                            if (use > 0) tc::mbar_wait(&bars.w_empty[s], (use - 1) & 1);
                            if (leader) {
                                tc::mbar_expect_tx(&bars.w_full[s], slot_stride);
                                tc::bulk_g2s_addr(w_ring + s * slot_stride, src + u * slot_stride, slot_stride,
                                                  &bars.w_full[s]);
                            }
                            prod++;
                        }
            }
    } else if (warp_u == 9) {
        // ------------------------------------------------------------ MMA issuer
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::make_idesc_tf32(TM, p.Npad);
        uint32_t cons = 0, acons = 0;
        int t_local = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += stride, t_local++) {
            const uint32_t tmem_d = tmem + (uint32_t)(t_local & 1) * (uint32_t)p.Npad;
            for (int a = 0; a < vatoms_per_tile; a++) {
                const uint32_t st = acons & (uint32_t)(p.tstages - 1);
                tc::mbar_wait(&bars.a_full[st], (acons / (uint32_t)p.tstages) & 1);
                const uint32_t ahi = tmem + 512u - 64u * (st + 1u), alo = ahi + 32u;
                {   // hi weights: A_hi.B_hi + A_lo.B_hi
                    const uint32_t s = cons & (WSLOTS - 1);
                    tc::mbar_wait(&bars.w_full[s], (cons / WSLOTS) & 1);
                    tc::tc_fence_after();
                    const uint64_t bd = tc::make_desc(w_ring + s * slot_stride);
                    if (leader) {
#pragma unroll
                        for (int k8 = 0; k8 < 4; k8++) {
                            tc::mma_tf32_ts(tmem_d, ahi + k8 * 8, bd + (uint64_t)(2 * k8), idesc,
                                            (a == 0 && k8 == 0) ? 0u : 1u);
                            tc::mma_tf32_ts(tmem_d, alo + k8 * 8, bd + (uint64_t)(2 * k8), idesc, 1u);
                        }
                        tc::mma_commit(&bars.w_empty[s]);
                    }
                    cons++;
                }
                {   // lo weights: A_hi.B_lo
                    const uint32_t s = cons & (WSLOTS - 1);
                    tc::mbar_wait(&bars.w_full[s], (cons / WSLOTS) & 1);
                    tc::tc_fence_after();
                    const uint64_t bd = tc::make_desc(w_ring + s * slot_stride);
                    if (leader) {
#pragma unroll
                        for (int k8 = 0; k8 < 4; k8++)
                            tc::mma_tf32_ts(tmem_d, ahi + k8 * 8, bd + (uint64_t)(2 * k8), idesc, 1u);
                        tc::mma_commit(&bars.w_empty[s]);
                        tc::mma_commit(&bars.a_free[st]);
                    }
                    cons++;
                }
                acons++;
            }
            if (leader) tc::mma_commit(&bars.d_full[t_local & 1]);
        }
    } else {
        // ------------------------------------------------------------ workers
        const int warp = tid >> 5, lane = tid & 31;
        const int row = 32 * (warp & 3) + lane;          // TMEM lane / tile row owned by this thread
        const int half = warp >> 2;                      // which 16 of an atom's 32 columns
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;

        // Epilogue: accumulator row -> bias / activation in the thread that owns the row, then through
        // a per-warp 4 KB shared-memory block (XOR-swizzled chunks, warp-local: __syncwarp only) so
        // that the global stores (and the skip loads) are whole 128-byte row segments: 8 lanes per
        // row, 4 rows per instruction, instead of 32 scattered 16-byte pieces.
        unsigned char *ostage = base + ASTAGES * A_STAGE_BYTES + WSLOTS * slot_stride + 256 + (size_t)warp * 4096;
        auto epilogue = [&](int tile, int t_local) {
            tc::mbar_wait(&bars.d_full[t_local & 1], (t_local >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t tmem_d = tmem + (uint32_t)(t_local & 1) * (uint32_t)p.Npad;
            const int64_t row_base = (int64_t)tile * TM + 32 * (warp & 3);   // first row of this warp
            const bool vec_c = (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
            const bool vec_s = p.skip != nullptr && (p.ldskip & 3) == 0 &&
                               ((reinterpret_cast<uintptr_t>(p.skip) & 15) == 0);
            for (int c0 = half * 32; c0 < p.Npad; c0 += 64) {
                float v[32];
                tc::tmem_ld32(tmem_d + lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++) {   // own row -> staging (bias added here: column-wise)
                    const int c = c0 + 4 * j4;
                    float4 o = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                    if (p.bias != nullptr && c + 3 < p.N) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + c));
                        o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                    } else if (p.bias != nullptr) {
                        if (c < p.N) o.x += __ldg(p.bias + c);
                        if (c + 1 < p.N) o.y += __ldg(p.bias + c + 1);
                        if (c + 2 < p.N) o.z += __ldg(p.bias + c + 2);
                    }
                    *reinterpret_cast<float4 *>(ostage + lane * 128 + ((j4 ^ (lane & 7)) << 4)) = o;
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; j++) {      // 8 lanes per row, rows (lane >> 3) + 4 j
                    const int r = (lane >> 3) + 4 * j, ch = lane & 7;
                    const int c = c0 + 4 * ch;
                    const int64_t grow = row_base + r;
                    float4 o = *reinterpret_cast<const float4 *>(ostage + r * 128 + ((ch ^ (r & 7)) << 4));
                    if (grow < p.M && c < p.N) {
                        float ov[4] = {o.x, o.y, o.z, o.w};
                        if (p.skip != nullptr) {
                            const float *srow = p.skip + (size_t)grow * p.ldskip + c;
                            if (vec_s && c + 3 < p.N) {
                                const float4 s4 = __ldg(reinterpret_cast<const float4 *>(srow));
                                ov[0] += s4.x; ov[1] += s4.y; ov[2] += s4.z; ov[3] += s4.w;
                            } else {
#pragma unroll
                                for (int q = 0; q < 4; q++) if (c + q < p.N) ov[q] += __ldg(srow + q);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) ov[q] = act_apply_compact(p.act, ov[q]);
                        float *crow = p.C + (size_t)grow * p.ldc + c;
                        if (vec_c && c + 3 < p.N) {
                            *reinterpret_cast<float4 *>(crow) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; q++) if (c + q < p.N) crow[q] = ov[q];
                        }
                    }
                }
                __syncwarp();
            }
        };

        // prime the raw-A ring: ASTAGES - 1 atoms ahead
        AtomPos pf{(int)blockIdx.x, 0, 0};
        for (int i = 0; i < ASTAGES - 1; i++) {
            issue_atom(p, pf, a_ring + (uint32_t)(i % ASTAGES) * A_STAGE_BYTES, tid);
            pf.next(p, stride);
        }
        uint32_t g = 0;          // physical atoms consumed so far (position in the shared-memory ring)
        uint32_t vg = 0;         // hand-offs to the MMAs so far (tensor-memory stage / barrier phase)
        int t_local = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += stride, t_local++) {
            float amp = 1.0f, att = 1.0f;   // PNA degree scalers of this thread's row (lib:1973-1984)
            if (p.rep1 == 3) {
                const int64_t grow = (int64_t)tile * TM + row;
                const int dg = grow < p.M ? __ldg(p.expand_deg + grow) : 1;
                const float lg1 = logf((float)((dg < 1 ? 1 : dg) + 1));
                amp = __fdiv_rn(lg1, p.expand_delta);
                att = __fdiv_rn(p.expand_delta, lg1);
            }
            for (int a = 0; a < atoms_per_tile; a++, g++) {
                cp_async_wait<ASTAGES - 2>();       // this thread's copies of atom g have landed
                worker_sync();                      // everyone's have, and atom g-1 has been read by all
                issue_atom(p, pf, a_ring + (uint32_t)((g + ASTAGES - 1) % ASTAGES) * A_STAGE_BYTES, tid);
                pf.next(p, stride);
                // own row, columns [16 half, +16) of the atom -> hi / lo -> tensor memory stage
                const unsigned char *stg = base + (size_t)(g % ASTAGES) * A_STAGE_BYTES + (size_t)row * tc::ROW_BYTES;
                float vv[16];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int chunk = half * 4 + q;
                    const float4 v = *reinterpret_cast<const float4 *>(stg + ((chunk ^ (row & 7)) << 4));
                    vv[4 * q] = v.x; vv[4 * q + 1] = v.y; vv[4 * q + 2] = v.z; vv[4 * q + 3] = v.w;
                }
                const int reps = a >= p.KA[0] ? p.rep1 : 1;
                for (int r = 0; r < reps; r++, vg++) {
                    const float sc = r == 0 ? 1.0f : (r == 1 ? amp : att);
                    float h[16], l[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float t = vv[j] * sc;
                        h[j] = tc::tf32_hi(t);
                        l[j] = t - h[j];
                    }
                    const uint32_t st = vg & (uint32_t)(p.tstages - 1), use = vg / (uint32_t)p.tstages;
                    if (use > 0) tc::mbar_wait(&bars.a_free[st], (use - 1) & 1);   // MMAs two hand-offs back are done
                    tc::tc_fence_after();
                    const uint32_t dst = tmem + 512u - 64u * (st + 1u) + lane_base + (uint32_t)(half * 16);
                    tc::tmem_st16(dst, h);
                    tc::tmem_st16(dst + 32u, l);
                    tc::tmem_st_wait();
                    tc::tc_fence_before();
                    mbar_arrive(&bars.a_full[st]);
                }
                // the previous tile's accumulator is complete by now: write it out while this
                // tile's MMAs run
                if (a == 0 && t_local > 0) epilogue(tile - stride, t_local - 1);
            }
        }
        if (t_local > 0) epilogue((int)blockIdx.x + (t_local - 1) * stride, t_local - 1);
        cp_async_wait<0>();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp_u == 0) tc::tmem_dealloc(tmem, 512);
}

// Wt[k][ldw] (transposed packed weights, device) -> per-atom [hi Npad x 128 B | lo Npad x 128 B]
__global__ void build_image_kernel(const float *__restrict__ Wt, int ldw, int K, int N, int Npad,
                                   int KA, float *__restrict__ img)
{
    const int64_t total = (int64_t)KA * Npad * tc::ATOM_K;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(i % tc::ATOM_K);
        const int n = (int)((i / tc::ATOM_K) % Npad);
        const int ka = (int)(i / ((int64_t)tc::ATOM_K * Npad));
        const int k = ka * tc::ATOM_K + kk;
        const float v = (k < K && n < N) ? __ldg(Wt + (size_t)k * ldw + n) : 0.0f;
        const float h = tc::tf32_hi(v);
        float *hi = img + (size_t)ka * 2 * Npad * tc::ATOM_K;
        float *lo = hi + (size_t)Npad * tc::ATOM_K;
        const uint32_t off = tc::canon_offset(n, kk, Npad) / 4;
        hi[off] = h;
        lo[off] = v - h;
    }
}

}  // namespace

size_t gemm_tc_image_floats(int K, int N)
{
    const int KA = (K + tc::ATOM_K - 1) / tc::ATOM_K, Npad = (N + 15) / 16 * 16;
    return (size_t)KA * 2 * Npad * tc::ATOM_K;
}

int gemm_tc_build_image(const float *Wt, int ldw, int K, int N, float *img, cudaStream_t s)
{
    const int KA = (K + tc::ATOM_K - 1) / tc::ATOM_K, Npad = (N + 15) / 16 * 16;
    const int64_t total = (int64_t)KA * Npad * tc::ATOM_K;
    if (total <= 0) return GNNB_OK;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, kNumSMs * 8);
    build_image_kernel<<<grid, 256, 0, s>>>(Wt, ldw, K, N, Npad, KA, img);
    GNNB_CUDA(cudaGetLastError());
    return GNNB_OK;
}

bool gemm_tc_supported(const GemmArgs &g)
{
    static const bool disabled = getenv("GNNB_DISABLE_TC_GEMM") != nullptr;
    if (disabled || g.img1 == nullptr || (g.A2 != nullptr && g.img2 == nullptr)) return false;
    const int Npad = (g.N + 15) / 16 * 16;
    // two accumulators + the A stages must fit the 512 columns; tiny problems are not worth a
    // persistent launch
    if (g.expand_deg != nullptr && (g.A2 == nullptr || g.K2 % (3 * tc::ATOM_K) != 0)) return false;
    return g.N >= 8 && 2 * Npad <= 384 && g.M >= 4 * TM && g.K1 >= 1;
}

int launch_gemm_tc(const GemmArgs &g, cudaStream_t s, int *launches)
{
    TcGemmParams p{};
    p.A[0] = g.A1; p.lda[0] = g.lda1; p.K[0] = g.K1; p.KA[0] = (g.K1 + tc::ATOM_K - 1) / tc::ATOM_K;
    p.img[0] = g.img1;
    if (g.A2 != nullptr && g.K2 > 0) {
        p.A[1] = g.A2; p.lda[1] = g.lda2; p.K[1] = g.K2; p.KA[1] = (g.K2 + tc::ATOM_K - 1) / tc::ATOM_K;
        p.img[1] = g.img2;
    }
    p.bias = g.bias; p.skip = g.skip; p.ldskip = g.ldskip; p.act = g.act;
    p.C = g.C; p.ldc = g.ldc; p.M = g.M; p.N = g.N; p.Npad = (g.N + 15) / 16 * 16;
    p.n_tiles = (g.M + TM - 1) / TM;
    // 4 stages (possible when 2 Npad <= 256) measured no faster than 2: the hand-off is not the limiter
    p.tstages = (getenv("GNNB_TC_GEMM_STAGES4") != nullptr && 2 * p.Npad <= 256) ? 4 : 2;
    p.rep1 = 1;
    if (g.expand_deg != nullptr && g.A2 != nullptr) {   // A2 = [M][K2/3] statistics, expanded x3
        p.K[1] = g.K2 / 3; p.KA[1] = p.K[1] / tc::ATOM_K; p.rep1 = 3;
        p.expand_deg = g.expand_deg; p.expand_delta = g.expand_delta;
    }
    const size_t smem = 1024 + (size_t)ASTAGES * A_STAGE_BYTES + (size_t)WSLOTS * p.Npad * tc::ROW_BYTES +
                        256 /* Bars */ + 8 * 4096 /* per-warp output staging */;
    static_assert(sizeof(Bars) <= 256, "Bars must fit its slot");
    GNNB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
    gemm_tc_kernel<<<grid, NTHREADS, smem, s>>>(p);
    GNNB_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    return GNNB_OK;
}

}  // namespace gnnb
