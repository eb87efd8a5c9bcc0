// Kernel (1), tensor-core version: the persistent whole-model fused kernel of fused.cu with the
// node-transform GEMMs on the 5th-generation tensor cores (tcgen05.mma.kind::tf32, accumulators in
// tensor memory) using the error-compensated 3xTF32 split of tc.cuh, so results stay fp32-grade
// (<= 1e-4 parity bound with two orders of magnitude of margin).
//
// Per CTA (one per SM, 256 threads, 128-row tile of packed graphs):
//   shared memory   R_X  64 KB  layer input X, canonical K-major SWIZZLE_128B layout; while a GEMM
//                               runs it is the ring of weight slots the bulk copies land in
//                   R_HI 64 KB  A operand, hi parts   (aggregate, then the GIN hidden layer)
//                   R_LO 64 KB  A operand, lo parts
//   tensor memory   128 columns x 128 lanes fp32 accumulator
//   weights         pre-split (hi/lo) and pre-swizzled on the host into the exact shared-memory
//                   image of each 32-wide K atom, so one cp.async.bulk (TMA engine, no tensor map)
//                   per atom lands them ready for the MMA; they stay L2 resident across CTAs.
// One elected thread issues the bulk copies and the MMAs and signals completion through mbarriers.
// A GEMM is a sequence of "units": first the hi weight atoms (each: A_hi.B_hi and A_lo.B_hi, four
// k-steps of 8), then the lo atoms (A_hi.B_lo).  With N = 128 the ring holds 4 slots of 16 KB, so
// ALL hi atoms of a K = 128 GEMM are in flight at once (one L2 latency), and each lo atom is
// fetched into the slot of a finished hi unit while later units run.  GIN's second GEMM has its
// first slots prefetched while the first GEMM's epilogue runs.  All 8 warps run the aggregation
// before and the TMEM -> register -> shared epilogue (bias, skip, activation, hi/lo re-split) after.
// The skip connection of interior layers needs X after R_X has been recycled for the weight ring:
// X is parked in a per-CTA global scratch (64 KB, L2 resident) and re-read in the epilogue.
// The MLP head (fp32 FMA, 16 graphs per call) is deferred: pooled vectors queue in a per-CTA
// pending buffer and the head runs once 16 graphs are waiting, so its weight streaming is
// amortised over ~3 tiles.
//
// Supported: GCN and GIN, every layer width a multiple of 16 up to 128 (BASELINE configs 1 and 2);
// everything else uses fused.cu / the layerwise path.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "model.h"
#include "tc.cuh"

namespace gnnb {

void build_weight_image(const float *W, int N, int n_valid, int K, int ld, int col0,
                        std::vector<float> &img);

namespace {

constexpr int TM = 128;
constexpr int ECAP = 2048;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int MAX_LAYERS = 8;
constexpr int MAX_HEAD = 6;
constexpr int HEAD_G = 128;      // pooled graphs per head call (one 128-row MMA tile)
constexpr int MAX_HCHUNK = 4;    // 128-wide K chunks of the first head layer (head_in <= 512)
constexpr int MAX_DIM = 128;
constexpr int PLD = 520;         // row stride of the pending pooled vectors (>= 512 + 4)
constexpr int REGION = TM * MAX_DIM * 4;  // 64 KB
constexpr int MAX_SLOTS = 8;
constexpr int MAX_NODES_PER_GRAPH = 64;
constexpr int AGG_R = 4;         // destination rows aggregated concurrently per warp

struct TLinear {
    const float *img;   // weight image: per K atom [hi N x 128 B | lo N x 128 B]
    const float *bias;
    int K, N, KA;
};

struct TcParams {
    int conv_type, num_layers, in_dim, skip, gnn_act, num_pools, pools[4];
    int mlp_num_linear, mlp_act, out_act, emb, mlp_out;
    float gin_eps;
    int fi[MAX_LAYERS], fo[MAX_LAYERS];
    TLinear l0[MAX_LAYERS], l1[MAX_LAYERS];
    TLinear hl[MAX_HEAD][MAX_HCHUNK];   // MLP head linears, K cut into 128-wide chunks
    int hchunks[MAX_HEAD];
    int head_n[MAX_HEAD];               // true output widths (N of the images is padded to 16)
    const float *x;
    const int32_t *coo;
    const int64_t *node_ptr, *edge_ptr;
    int n_graphs;
    float *out;
    const int32_t *tile_bounds;
    int n_tiles;
    int *error_flag;
    float *scratch;               // [grid][128][128] parked X for the skip connection
    float *pending;               // [grid][HEAD_G][PLD] pooled vectors waiting for the head
    unsigned long long *timing;
};

struct Misc {
    uint64_t bar_full[MAX_SLOTS], bar_empty[MAX_SLOTS], bar_done;
    uint32_t full_cnt[MAX_SLOTS], empty_cnt[MAX_SLOTS];
    uint32_t tmem_slot;
    int pend_n;
    int pend_gid[HEAD_G];
    float dinv[TM];
    int deg[TM];
    int off[TM + 1];
    int rowg[TM];
    int grow[TM + 2];
    int gedge[TM + 2];
    int scan_tmp[8];
    unsigned short edges[ECAP];
    unsigned char nbr[ECAP];
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ int lower_bound64(const int64_t *__restrict__ ptr, int n, int64_t v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void tc_tile_bounds_kernel(const int64_t *__restrict__ node_ptr, int n_graphs,
                                      int window, int n_tiles, int32_t *__restrict__ bounds)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    bounds[t] = (t == n_tiles) ? n_graphs : lower_bound64(node_ptr, n_graphs, (int64_t)t * window);
}

// ---------------------------------------------------------------------------------------
// GEMM issue (ONE thread).  D[128][N] (tensor memory) = A(hi,lo)[128][KA*32] . W^T, 3xTF32.
__device__ __forceinline__ int gemm_slots(const TLinear &L)
{
    const int s = REGION / (L.N * tc::ROW_BYTES);
    return s < MAX_SLOTS ? s : MAX_SLOTS;
}
__device__ __forceinline__ const unsigned char *unit_src(const TLinear &L, int u)
{
    const int part = u >= L.KA ? 1 : 0, ka = part ? u - L.KA : u;
    return reinterpret_cast<const unsigned char *>(L.img) +
           ((size_t)ka * 2 + part) * (size_t)L.N * tc::ROW_BYTES;
}
// start the bulk copies of the first min(slots, units) weight slots
__device__ __forceinline__ void gemm_prefetch(Misc &ms, unsigned char *ring, const TLinear &L)
{
    const uint32_t slot_bytes = (uint32_t)L.N * tc::ROW_BYTES;
    const int U = 2 * L.KA, ns = gemm_slots(L);
    for (int u = 0; u < ns && u < U; u++) {
        tc::mbar_expect_tx(&ms.bar_full[u], slot_bytes);
        tc::bulk_g2s(ring + (size_t)u * slot_bytes, unit_src(L, u), slot_bytes, &ms.bar_full[u]);
    }
}
__device__ __forceinline__ void gemm_issue(Misc &ms, uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo,
                                           unsigned char *ring, const TLinear &L, bool prefetched,
                                           bool accumulate = false)
{
    const uint32_t slot_bytes = (uint32_t)L.N * tc::ROW_BYTES;
    const uint32_t idesc = tc::make_idesc_tf32(TM, L.N);
    const int U = 2 * L.KA, ns = gemm_slots(L);
    if (!prefetched) gemm_prefetch(ms, ring, L);
    tc::tc_fence_after();
    for (int u = 0; u < U; u++) {
        const int s = u % ns;
        tc::mbar_wait(&ms.bar_full[s], ms.full_cnt[s] & 1);
        ms.full_cnt[s]++;
        tc::tc_fence_after();
        const int part = u >= L.KA ? 1 : 0, ka = part ? u - L.KA : u;
        const uint32_t ah = a_hi + (uint32_t)ka * TM * tc::ROW_BYTES;
        const uint32_t al = a_lo + (uint32_t)ka * TM * tc::ROW_BYTES;
        const uint32_t b = tc::smem_u32(ring) + (uint32_t)s * slot_bytes;
#pragma unroll
        for (int k8 = 0; k8 < tc::ATOM_K / tc::MMA_K; k8++) {
            const uint32_t ko = (uint32_t)k8 * tc::MMA_K * 4;
            if (part == 0) {
                tc::mma_tf32(tmem_d, tc::make_desc(ah + ko), tc::make_desc(b + ko), idesc,
                             (u == 0 && k8 == 0 && !accumulate) ? 0u : 1u);
                tc::mma_tf32(tmem_d, tc::make_desc(al + ko), tc::make_desc(b + ko), idesc, 1u);
            } else {
                tc::mma_tf32(tmem_d, tc::make_desc(ah + ko), tc::make_desc(b + ko), idesc, 1u);
            }
        }
        tc::mma_commit(&ms.bar_empty[s]);
        ms.empty_cnt[s]++;
        // refill the slot of the PREVIOUS unit (its MMAs were issued before this unit's, so the
        // wait is short and this unit's MMAs keep the tensor core busy meanwhile)
        if (u >= 1 && u - 1 + ns < U) {
            const int sp = (u - 1) % ns;
            tc::mbar_wait(&ms.bar_empty[sp], (ms.empty_cnt[sp] - 1) & 1);
            tc::mbar_expect_tx(&ms.bar_full[sp], slot_bytes);
            tc::bulk_g2s(ring + (size_t)sp * slot_bytes, unit_src(L, u - 1 + ns), slot_bytes,
                         &ms.bar_full[sp]);
        }
    }
    tc::mma_commit(&ms.bar_done);
}

// All threads: accumulator -> (+bias, +skip, activation) -> destination.
//   EPI_X      : write the values into dst_hi (the X region, canonical layout)
//   EPI_SPLIT  : write hi / lo parts into dst_hi / dst_lo (the next GEMM's A operand)
//   EPI_GLOBAL : write columns [0, n_true) of rows [0, n_rows) to gout[gids[row]][col] (model output)
// Columns [N, round_up(N, 32)) of the last K atom are zero filled (shared-memory modes).
// `skip` is written by this CTA earlier in the same kernel: read it with ld.global.cg (L2), never
// through the non-coherent read-only path.
enum { EPI_X = 0, EPI_SPLIT = 1, EPI_GLOBAL = 2 };
__device__ __forceinline__ void epilogue(uint32_t tmem_d, int N, const float *__restrict__ bias,
                                         int act, const float *skip, int mode,
                                         unsigned char *dst_hi, unsigned char *dst_lo,
                                         float *gout = nullptr, const int *gids = nullptr,
                                         int n_rows = 0, int ldg = 0, int n_true = 0)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = 32 * (warp & 3) + lane;
    const int npad = (N + 31) & ~31;
    float *grow = nullptr;
    if (mode == EPI_GLOBAL && row < n_rows) grow = gout + (size_t)gids[row] * ldg;
    for (int c0 = (warp >> 2) * 32; c0 < npad; c0 += 64) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
        // bias (warp-uniform addresses) and the parked skip rows: issue all loads of this block
        // up front so that their latencies overlap
        float4 bsv[8], skv[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
            const bool in_range = c0 + j4 * 4 < N;  // N % 4 == 0
            bsv[j4] = in_range ? __ldg(reinterpret_cast<const float4 *>(bias + c0 + j4 * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            skv[j4] = (in_range && skip != nullptr)
                          ? __ldcg(reinterpret_cast<const float4 *>(skip + (size_t)row * MAX_DIM + c0 + j4 * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
            float o[4];
            const float4 sk = skv[j4], bs = bsv[j4];
            const bool in_range = c0 + j4 * 4 < N;
            const float sks[4] = {sk.x, sk.y, sk.z, sk.w};
            const float bss[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
            for (int j = 0; j < 4; j++)
                o[j] = in_range ? act_apply_compact(act, v[j4 * 4 + j] + bss[j] + sks[j]) : 0.0f;
            if (mode == EPI_GLOBAL) {
                if (grow != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (c0 + j4 * 4 + j < n_true) grow[c0 + j4 * 4 + j] = o[j];
                }
                continue;
            }
            const uint32_t off = tc::canon_chunk_offset(row, c0 + j4 * 4, TM);
            if (mode == EPI_SPLIT) {
                float h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { h[j] = tc::tf32_hi(o[j]); l[j] = o[j] - h[j]; }
                *reinterpret_cast<float4 *>(dst_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4 *>(dst_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
            } else {
                *reinterpret_cast<float4 *>(dst_hi + off) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// MLP head (cpp:454-530) for up to 128 pending graphs on the tensor cores: the pooled vectors
// [128][head_in] are the A operand, read from the per-CTA pending buffer in 128-wide K chunks that
// accumulate into the same TMEM tile; later head layers take their A operand from the previous
// epilogue exactly like GIN's hidden layer.  All threads call this.
__device__ __forceinline__ void head_flush(const TcParams &p, Misc &ms, uint32_t tmem_d,
                                           unsigned char *RX, unsigned char *RHI, unsigned char *RLO,
                                           const float *pending, int n_rows, uint32_t &done_cnt)
{
    const int tid = threadIdx.x;
    const int head_in = p.emb * p.num_pools;
    for (int j = 0; j < p.mlp_num_linear; j++) {
        const bool last = j == p.mlp_num_linear - 1;
        for (int c = 0; c < p.hchunks[j]; c++) {
            const TLinear &L = p.hl[j][c];
            if (j == 0) {  // A chunk: pending[:, 128c : 128c + K) -> (hi, lo), zero padded
                const int kp = L.KA * tc::ATOM_K, q4 = kp / 4;
                for (int idx = tid; idx < TM * q4; idx += NTHREADS) {
                    const int r = idx / q4, cc = (idx - r * q4) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < n_rows && c * 128 + cc < head_in)  // head_in % 4 == 0
                        v = __ldcg(reinterpret_cast<const float4 *>(pending + (size_t)r * PLD + c * 128 + cc));
                    const float4 h = make_float4(tc::tf32_hi(v.x), tc::tf32_hi(v.y), tc::tf32_hi(v.z),
                                                 tc::tf32_hi(v.w));
                    const uint32_t off = tc::canon_chunk_offset(r, cc, TM);
                    *reinterpret_cast<float4 *>(RHI + off) = h;
                    *reinterpret_cast<float4 *>(RLO + off) =
                        make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                }
                tc::fence_async_smem();
                tc::tc_fence_before();
                __syncthreads();
            }
            if (tid == 0) {
                gemm_issue(ms, tmem_d, tc::smem_u32(RHI), tc::smem_u32(RLO), RX, L, false, c > 0);
                tc::mbar_wait(&ms.bar_done, done_cnt & 1);
            }
            done_cnt++;
            __syncthreads();
            tc::tc_fence_after();
        }
        const TLinear &L0 = p.hl[j][0];
        if (last)
            epilogue(tmem_d, L0.N, L0.bias, p.out_act, nullptr, EPI_GLOBAL, nullptr, nullptr, p.out,
                     ms.pend_gid, n_rows, p.mlp_out, p.head_n[j]);
        else
            epilogue(tmem_d, L0.N, L0.bias, p.mlp_act, nullptr, EPI_SPLIT, RHI, RLO);
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) fused_tc_kernel(const TcParams p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *RX = base, *RHI = base + REGION, *RLO = base + 2 * REGION;
    Misc &ms = *reinterpret_cast<Misc *>(base + 3 * REGION);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_prev = clock64();
#define GNNB_PHASE(idx)                                                             \
    if (p.timing != nullptr && tid == 0) {                                          \
        const long long t_now = clock64();                                          \
        atomicAdd(p.timing + (idx), (unsigned long long)(t_now - t_prev));          \
        t_prev = t_now;                                                             \
    }

    if (warp == 0) tc::tmem_alloc(&ms.tmem_slot, 128);
    if (tid == 0) {
        for (int i = 0; i < MAX_SLOTS; i++) {
            tc::mbar_init(&ms.bar_full[i], 1);
            tc::mbar_init(&ms.bar_empty[i], 1);
            ms.full_cnt[i] = 0;
            ms.empty_cnt[i] = 0;
        }
        tc::mbar_init(&ms.bar_done, 1);
        tc::mbar_fence_init();
        ms.pend_n = 0;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = ms.tmem_slot;
    uint32_t done_cnt = 0;
    float *scratch = p.scratch + (size_t)blockIdx.x * TM * MAX_DIM;
    float *pending = p.pending + (size_t)blockIdx.x * HEAD_G * PLD;
    const int emb = p.emb;

    // geometry of the first tile; later tiles are prefetched one iteration ahead
    int g0 = 0, g1 = 0;
    int64_t row0 = 0, e0 = 0, row1 = 0, e1 = 0;
    if ((int)blockIdx.x < p.n_tiles) {
        g0 = __ldg(p.tile_bounds + blockIdx.x);
        g1 = __ldg(p.tile_bounds + blockIdx.x + 1);
        row0 = __ldg(p.node_ptr + g0); row1 = __ldg(p.node_ptr + g1);
        e0 = __ldg(p.edge_ptr + g0); e1 = __ldg(p.edge_ptr + g1);
    }

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int ng = g1 - g0;
        const int rows = (int)(row1 - row0);
        const int ne = (int)(e1 - e0);
        const int cur_g0 = g0;
        const int64_t cur_row0 = row0, cur_e0 = e0;
        // next tile's graph range: issue the loads now, consume them after the table build
        const int nxt = tile + gridDim.x;
        int ng0 = 0, ng1 = 0;
        if (nxt < p.n_tiles) {
            ng0 = __ldg(p.tile_bounds + nxt);
            ng1 = __ldg(p.tile_bounds + nxt + 1);
        }
        const bool bad = rows > TM || ne > ECAP || ng > TM;
        if (bad && tid == 0) atomicExch(p.error_flag, 1);
        if (ng > 0 && !bad) {
            __syncthreads();
            // ---------------------------------------------------------------- stage inputs
            for (int i = tid; i <= ng; i += NTHREADS) {
                ms.grow[i] = (int)(__ldg(p.node_ptr + cur_g0 + i) - cur_row0);
                ms.gedge[i] = (int)(__ldg(p.edge_ptr + cur_g0 + i) - cur_e0);
            }
            {   // node features -> R_X (canonical layout), zero padded to whole K atoms; rows
                // beyond the tile are zeroed too so no stale NaN/Inf ever enters an MMA
                const int F = p.in_dim, Fp = (F + 31) & ~31;
                const float *src = p.x + (size_t)cur_row0 * F;
                // batches of 8 independent loads per thread so the global latency is paid once per
                // batch, not once per element
                for (int b0 = 0; b0 < TM * Fp; b0 += NTHREADS * 8) {
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int idx = b0 + q * NTHREADS + tid;
                        const int r = idx / Fp, c = idx - r * Fp;
                        v[q] = (r < rows && c < F) ? __ldg(src + (size_t)r * F + c) : 0.0f;
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const int idx = b0 + q * NTHREADS + tid;
                        const int r = idx / Fp, c = idx - r * Fp;
                        if (r < TM) *reinterpret_cast<float *>(RX + tc::canon_offset(r, c, TM)) = v[q];
                    }
                }
            }
            int2 my_edge[ECAP / NTHREADS];
            {   // edge list: issue the global loads before the barrier, localise after it
                const int2 *coo = reinterpret_cast<const int2 *>(p.coo) + cur_e0;
#pragma unroll
                for (int q = 0; q < ECAP / NTHREADS; q++) {
                    const int j = tid + q * NTHREADS;
                    my_edge[q] = (j < ne) ? __ldg(coo + j) : make_int2(0, 0);
                }
            }
            __syncthreads();
            for (int i = tid; i < ng; i += NTHREADS)
                for (int r = ms.grow[i]; r < ms.grow[i + 1]; r++) ms.rowg[r] = i;
#pragma unroll
            for (int q = 0; q < ECAP / NTHREADS; q++) {
                const int j = tid + q * NTHREADS;
                if (j < ne) {
                    const int2 sd = my_edge[q];
                    int lo = 0, hi = ng;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (ms.gedge[mid] <= j) lo = mid; else hi = mid;
                    }
                    const int b = ms.grow[lo], n_i = ms.grow[lo + 1] - b;
                    if ((unsigned)sd.x >= (unsigned)n_i || (unsigned)sd.y >= (unsigned)n_i) {
                        atomicExch(p.error_flag, 2);
                        ms.edges[j] = 0xffff;
                    } else {
                        ms.edges[j] = (unsigned short)(((sd.y + b) << 8) | (sd.x + b));
                    }
                }
            }
            __syncthreads();
            GNNB_PHASE(0)
            // ---------------------------------------------------------------- tables (lib:1051-1124)
            int my_deg = 0;
            if (tid < TM) {
                if (tid < rows) {
                    const int gi = ms.rowg[tid];
                    for (int j = ms.gedge[gi]; j < ms.gedge[gi + 1]; j++)
                        my_deg += ((ms.edges[j] >> 8) == tid) ? 1 : 0;
                }
                ms.deg[tid] = my_deg;
                ms.dinv[tid] = 1.0f / sqrtf(1.0f + (float)my_deg);
                int incl = my_deg;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                if (lane == 31) ms.scan_tmp[warp] = incl;
                ms.off[tid] = incl - my_deg;
            }
            __syncthreads();
            if (tid < TM) {
                int b = 0;
                for (int w = 0; w < warp; w++) b += ms.scan_tmp[w];
                const int o = ms.off[tid] + b;
                ms.off[tid] = o;
                if (tid < rows) {
                    const int gi = ms.rowg[tid];
                    int pos = o;
                    for (int j = ms.gedge[gi]; j < ms.gedge[gi + 1]; j++) {
                        const unsigned short ed = ms.edges[j];
                        if ((ed >> 8) == tid) ms.nbr[pos++] = (unsigned char)(ed & 0xff);
                    }
                }
            }
            __syncthreads();
            GNNB_PHASE(1)
        }
        // second half of the next tile's geometry (the bounds have arrived by now)
        if (nxt < p.n_tiles) {
            g0 = ng0; g1 = ng1;
            row0 = __ldg(p.node_ptr + ng0); row1 = __ldg(p.node_ptr + ng1);
            e0 = __ldg(p.edge_ptr + ng0); e1 = __ldg(p.edge_ptr + ng1);
        }
        if (ng <= 0 || bad) continue;

        // -------------------------------------------------------------------- conv layers
        for (int l = 0; l < p.num_layers; l++) {
            const int fi = p.fi[l];
            const int kp = (fi + 31) & ~31;
            const bool do_skip = p.skip && l != 0 && l != p.num_layers - 1;  // cpp:269-279
            // aggregate X -> (hi, lo) A operand: a warp works on AGG_R rows at a time (independent
            // gather chains), lanes across K (float4 each)
            const int c = lane * 4;
            if (c < kp) {
                for (int rb = warp * AGG_R; rb < TM; rb += NWARPS * AGG_R) {
                    int d[AGG_R], o[AGG_R];
                    float4 acc[AGG_R];
                    int kmax = 0;
#pragma unroll
                    for (int j = 0; j < AGG_R; j++) {
                        const int r = rb + j;
                        d[j] = r < rows ? ms.deg[r] : 0;
                        o[j] = ms.off[r];
                        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        kmax = max(kmax, d[j]);
                    }
                    // branch-free body: all AGG_R gather chains (index -> row -> add) issue back to
                    // back; rows that have run out of neighbors read slot 0 and select 0
                    const bool gcn = p.conv_type == GNNB_CONV_GCN;
                    for (int k = 0; k < kmax; k++) {
                        int u[AGG_R];
                        float4 v[AGG_R];
                        float s[AGG_R];
#pragma unroll
                        for (int j = 0; j < AGG_R; j++) u[j] = ms.nbr[(k < d[j]) ? o[j] + k : 0];
#pragma unroll
                        for (int j = 0; j < AGG_R; j++) {
                            v[j] = *reinterpret_cast<const float4 *>(
                                RX + tc::canon_chunk_offset(u[j], c, TM));
                            s[j] = gcn ? ms.dinv[u[j]] : 1.0f;
                        }
#pragma unroll
                        for (int j = 0; j < AGG_R; j++) {
                            const bool on = k < d[j];
                            const float4 t = make_float4(on ? v[j].x : 0.0f, on ? v[j].y : 0.0f,
                                                         on ? v[j].z : 0.0f, on ? v[j].w : 0.0f);
                            if (gcn) {
                                acc[j].x = fmaf(t.x, s[j], acc[j].x); acc[j].y = fmaf(t.y, s[j], acc[j].y);
                                acc[j].z = fmaf(t.z, s[j], acc[j].z); acc[j].w = fmaf(t.w, s[j], acc[j].w);
                            } else {
                                acc[j].x += t.x; acc[j].y += t.y; acc[j].z += t.z; acc[j].w += t.w;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < AGG_R; j++) {
                        const int r = rb + j;
                        float4 a = acc[j];
                        if (r < rows) {
                            const float4 xs = *reinterpret_cast<const float4 *>(
                                RX + tc::canon_chunk_offset(r, c, TM));
                            if (p.conv_type == GNNB_CONV_GCN) {  // lib:1249-1278, factorised
                                const float dv = ms.dinv[r], ss = dv * dv;
                                a.x = fmaf(xs.x, ss, a.x * dv); a.y = fmaf(xs.y, ss, a.y * dv);
                                a.z = fmaf(xs.z, ss, a.z * dv); a.w = fmaf(xs.w, ss, a.w * dv);
                            } else {  // GIN, lib:1519-1529
                                const float s = 1.0f + p.gin_eps;
                                a.x += xs.x * s; a.y += xs.y * s; a.z += xs.z * s; a.w += xs.w * s;
                            }
                        }
                        const float4 h = make_float4(tc::tf32_hi(a.x), tc::tf32_hi(a.y),
                                                     tc::tf32_hi(a.z), tc::tf32_hi(a.w));
                        const uint32_t off = tc::canon_chunk_offset(r, c, TM);
                        *reinterpret_cast<float4 *>(RHI + off) = h;
                        *reinterpret_cast<float4 *>(RLO + off) =
                            make_float4(a.x - h.x, a.y - h.y, a.z - h.z, a.w - h.w);
                    }
                }
            }
            GNNB_PHASE(2)   // this warp's own aggregation time (no barrier yet)
            if (do_skip) {  // park X: R_X is about to become the weight ring
                for (int idx = tid; idx < TM * (kp / 4); idx += NTHREADS) {
                    const int r = idx / (kp / 4), cc = (idx % (kp / 4)) * 4;
                    const float4 v = *reinterpret_cast<const float4 *>(
                        RX + tc::canon_chunk_offset(r, cc, TM));
                    *reinterpret_cast<float4 *>(scratch + (size_t)r * MAX_DIM + cc) = v;
                }
            }
            tc::fence_async_smem();
            tc::tc_fence_before();
            __syncthreads();
            GNNB_PHASE(6)   // skip spill + proxy fence + barrier (waiting for the slowest warp)
            const float *skip = do_skip ? scratch : nullptr;
            if (tid == 0) {
                gemm_issue(ms, tmem_d, tc::smem_u32(RHI), tc::smem_u32(RLO), RX, p.l0[l], false);
                tc::mbar_wait(&ms.bar_done, done_cnt & 1);  // one poller; the rest park on bar.sync
            }
            done_cnt++;
            __syncthreads();
            tc::tc_fence_after();
            GNNB_PHASE(7)   // weight copies + MMAs until the accumulator is ready
            if (p.conv_type == GNNB_CONV_GCN) {
                epilogue(tmem_d, p.l0[l].N, p.l0[l].bias, p.gnn_act, skip, EPI_X, RX, nullptr);
            } else {
                // the ring is idle while the epilogue runs: start fetching the second GEMM's weights
                if (tid == 0) gemm_prefetch(ms, RX, p.l1[l]);
                epilogue(tmem_d, p.l0[l].N, p.l0[l].bias, GNNB_ACT_RELU, nullptr, EPI_SPLIT, RHI, RLO);
                tc::fence_async_smem();
                tc::tc_fence_before();
                __syncthreads();
                GNNB_PHASE(3)
                if (tid == 0) {
                    gemm_issue(ms, tmem_d, tc::smem_u32(RHI), tc::smem_u32(RLO), RX, p.l1[l], true);
                    tc::mbar_wait(&ms.bar_done, done_cnt & 1);
                }
                done_cnt++;
                __syncthreads();
                tc::tc_fence_after();
                GNNB_PHASE(7)
                epilogue(tmem_d, p.l1[l].N, p.l1[l].bias, p.gnn_act, skip, EPI_X, RX, nullptr);
            }
            tc::tc_fence_before();
            __syncthreads();
            GNNB_PHASE(3)
        }

        // -------------------------------------------------------------------- pooling -> pending
        for (int gdone = 0; gdone < ng;) {
            const int base_n = ms.pend_n;  // uniform: written below only after a barrier
            const int cnt = min(HEAD_G - base_n, ng - gdone);
            for (int gi = warp; gi < cnt; gi += NWARPS) {
                const int r0 = ms.grow[gdone + gi], r1 = ms.grow[gdone + gi + 1];
                float *dst = pending + (size_t)(base_n + gi) * PLD;
                for (int cc = lane * 4; cc < emb; cc += 128) {   // emb % 16 == 0
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), mx = sum;
                    for (int r = r0; r < r1; r++) {
                        const float4 v = *reinterpret_cast<const float4 *>(
                            RX + tc::canon_chunk_offset(r, cc, TM));
                        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                        const bool first = r == r0;                       // lib:748-759
                        mx.x = (first || v.x > mx.x) ? v.x : mx.x; mx.y = (first || v.y > mx.y) ? v.y : mx.y;
                        mx.z = (first || v.z > mx.z) ? v.z : mx.z; mx.w = (first || v.w > mx.w) ? v.w : mx.w;
                    }
                    const float inv_on = (r1 > r0) ? 1.0f : 0.0f, cntf = (float)max(r1 - r0, 1);
                    for (int q = 0; q < p.num_pools; q++) {
                        float4 v;
                        if (p.pools[q] == GNNB_POOL_ADD) v = sum;
                        else if (p.pools[q] == GNNB_POOL_MEAN)
                            v = make_float4(inv_on * (sum.x / cntf), inv_on * (sum.y / cntf),
                                            inv_on * (sum.z / cntf), inv_on * (sum.w / cntf));
                        else v = mx;
                        *reinterpret_cast<float4 *>(dst + q * emb + cc) = v;
                    }
                }
                if (lane == 0) ms.pend_gid[base_n + gi] = cur_g0 + gdone + gi;
            }
            __syncthreads();
            if (tid == 0) ms.pend_n = base_n + cnt;
            gdone += cnt;
            GNNB_PHASE(4)
            if (base_n + cnt == HEAD_G) {
                head_flush(p, ms, tmem_d, RX, RHI, RLO, pending, HEAD_G, done_cnt);
                if (tid == 0) ms.pend_n = 0;
                GNNB_PHASE(5)
            }
            __syncthreads();
        }
    }
    // graphs still waiting for the head
    __syncthreads();
    {
        const int left = ms.pend_n;
        if (left > 0) head_flush(p, ms, tmem_d, RX, RHI, RLO, pending, left, done_cnt);
    }
    GNNB_PHASE(5)
#undef GNNB_PHASE
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 128);
}

}  // namespace

struct TcPlan {
    TcParams params{};
    DeviceBuf images, bounds, flag, scratch, pending, timing;
    size_t smem_bytes = 0;
};

int fused_tc_prepare(gnnb_model *m)
{
    m->fused_tc = nullptr;
    const gnnb_model_desc &d = m->d;
    if (getenv("GNNB_DISABLE_TC") != nullptr) return GNNB_OK;
    const bool conv_ok = d.conv_type == GNNB_CONV_GCN || d.conv_type == GNNB_CONV_GIN;
    const int head_in = m->emb_dim() * d.num_pools;
    bool dims_ok = d.num_layers >= 1 && d.num_layers <= MAX_LAYERS && d.mlp_num_linear <= MAX_HEAD &&
                   d.in_dim <= MAX_DIM && d.mlp_hidden <= MAX_DIM && d.mlp_out <= MAX_DIM &&
                   head_in <= 512;
    for (int k = 0; k < d.num_layers && dims_ok; k++) {
        int fi, fo;
        m->layer_dims(k, &fi, &fo);
        dims_ok = fo % 16 == 0 && fo >= 16 && fo <= MAX_DIM && fi <= MAX_DIM;
    }
    if (!conv_ok || !dims_ok) return GNNB_OK;

    TcPlan *plan = new TcPlan();
    TcParams &p = plan->params;
    p.conv_type = d.conv_type; p.num_layers = d.num_layers; p.in_dim = d.in_dim; p.skip = d.skip;
    p.gnn_act = d.gnn_act; p.num_pools = d.num_pools;
    for (int i = 0; i < 4; i++) p.pools[i] = d.pools[i];
    p.mlp_num_linear = d.mlp_num_linear; p.mlp_act = d.mlp_act; p.out_act = d.out_act;
    p.emb = m->emb_dim(); p.mlp_out = d.mlp_out; p.gin_eps = d.gin_eps;

    // weight images from the host copies of the parameters (flat reference order: head first)
    std::vector<float> img;
    struct Pending { size_t off; int K, N; };
    std::vector<Pending> pend;
    size_t idx = 2 * (size_t)d.mlp_num_linear;
    for (int k = 0; k < d.num_layers; k++) {
        int fi, fo;
        m->layer_dims(k, &fi, &fo);
        if (d.conv_type == GNNB_CONV_GCN) {  // [bias, lin_weight]
            pend.push_back({img.size(), fi, fo});
            build_weight_image(m->params[idx + 1].host.data(), fo, fo, fi, fi, 0, img);
            idx += 2;
        } else {                              // [w0, b0, w1, b1]
            pend.push_back({img.size(), fi, fo});
            build_weight_image(m->params[idx].host.data(), fo, fo, fi, fi, 0, img);
            pend.push_back({img.size(), fo, fo});
            build_weight_image(m->params[idx + 2].host.data(), fo, fo, fo, fo, 0, img);
            idx += 4;
        }
    }
    // MLP head: N padded to a multiple of 16, K cut into 128-wide chunks, bias zero padded
    struct HeadPending { size_t off[MAX_HCHUNK]; int K[MAX_HCHUNK]; int nch, N, n_true; size_t bias; };
    std::vector<HeadPending> hpend;
    {
        int in = head_in;
        for (int j = 0; j < d.mlp_num_linear; j++) {
            const int out = (j == d.mlp_num_linear - 1) ? d.mlp_out : d.mlp_hidden;
            HeadPending h{};
            h.N = (out + 15) / 16 * 16;
            h.n_true = out;
            h.nch = (in + 127) / 128;
            const float *W = m->params[2 * (size_t)j].host.data();
            const float *b = m->params[2 * (size_t)j + 1].host.data();
            for (int c = 0; c < h.nch; c++) {
                h.K[c] = std::min(128, in - 128 * c);
                h.off[c] = img.size();
                build_weight_image(W, h.N, out, h.K[c], in, 128 * c, img);
            }
            h.bias = img.size();
            img.resize(img.size() + h.N, 0.0f);
            for (int i = 0; i < out; i++) img[h.bias + i] = b[i];
            hpend.push_back(h);
            in = out;
        }
    }
    int rc = plan->images.ensure(img.size() * sizeof(float));
    if (rc == GNNB_OK) {
        cudaError_t e = cudaMemcpy(plan->images.ptr, img.data(), img.size() * sizeof(float),
                                   cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = cuda_fail(e, "upload weight images", __FILE__, __LINE__);
    }
    if (rc == GNNB_OK) rc = plan->flag.ensure(sizeof(int));
    if (rc == GNNB_OK) rc = plan->scratch.ensure((size_t)kNumSMs * TM * MAX_DIM * sizeof(float));
    if (rc == GNNB_OK) rc = plan->pending.ensure((size_t)kNumSMs * HEAD_G * PLD * sizeof(float));
    if (rc == GNNB_OK && getenv("GNNB_FUSED_TIMING") != nullptr) {
        rc = plan->timing.ensure(16 * sizeof(unsigned long long));
        if (rc == GNNB_OK) cudaMemset(plan->timing.ptr, 0, 16 * sizeof(unsigned long long));
    }
    if (rc != GNNB_OK) { delete plan; return rc; }
    const float *base = plan->images.as<float>();
    size_t pi = 0;
    for (int k = 0; k < d.num_layers; k++) {
        const LayerPack &L = m->layers[k];
        p.fi[k] = L.fi; p.fo[k] = L.fo;
        auto mk = [&](const Pending &q, const float *bias) {
            TLinear t;
            t.img = base + q.off; t.bias = bias; t.K = q.K; t.N = q.N;
            t.KA = (q.K + tc::ATOM_K - 1) / tc::ATOM_K;
            return t;
        };
        p.l0[k] = mk(pend[pi++], L.a.bias);
        if (d.conv_type == GNNB_CONV_GIN) p.l1[k] = mk(pend[pi++], L.b.bias);
    }
    for (int j = 0; j < d.mlp_num_linear; j++) {
        const HeadPending &h = hpend[j];
        p.hchunks[j] = h.nch;
        p.head_n[j] = h.n_true;
        for (int c = 0; c < h.nch; c++) {
            TLinear t;
            t.img = base + h.off[c]; t.bias = base + h.bias; t.K = h.K[c]; t.N = h.N;
            t.KA = (h.K[c] + tc::ATOM_K - 1) / tc::ATOM_K;
            p.hl[j][c] = t;
        }
    }
    plan->smem_bytes = 1024 + 3 * (size_t)REGION + sizeof(Misc);
    cudaError_t e = cudaFuncSetAttribute(fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)plan->smem_bytes);
    if (e != cudaSuccess) {
        delete plan;
        return cuda_fail(e, "cudaFuncSetAttribute(fused_tc_kernel)", __FILE__, __LINE__);
    }
    m->fused_tc = plan;
    return GNNB_OK;
}

void fused_tc_release(gnnb_model *m)
{
    TcPlan *plan = m->fused_tc;
    if (plan) {
        plan->images.release(); plan->bounds.release(); plan->flag.release();
        plan->scratch.release(); plan->pending.release(); plan->timing.release();
        delete plan;
        m->fused_tc = nullptr;
    }
}

bool fused_tc_supports(const gnnb_model *m, int max_nodes_in_batch, int max_edges_in_batch)
{
    return m->fused_tc != nullptr && max_nodes_in_batch <= MAX_NODES_PER_GRAPH &&
           max_edges_in_batch <= ECAP / 4;
}

int fused_tc_run(gnnb_model *m, const float *x, const int32_t *coo, const int64_t *node_ptr,
                 const int64_t *edge_ptr, int n_graphs, int64_t total_nodes, int max_nodes,
                 float *out, cudaStream_t s, int *launches)
{
    TcPlan *plan = m->fused_tc;
    GNNB_REQUIRE(plan != nullptr, "tensor-core fused kernel not available for this model");
    if (max_nodes < 1) max_nodes = 1;
    GNNB_REQUIRE(max_nodes <= MAX_NODES_PER_GRAPH, "graph too large for the fused kernel");
    const int window = TM - max_nodes + 1;
    const int64_t n_tiles64 = total_nodes / window + 1;
    GNNB_REQUIRE(n_tiles64 < (1ll << 30), "too many tiles");
    const int n_tiles = (int)n_tiles64;
    GNNB_TRY(plan->bounds.ensure(sizeof(int32_t) * ((size_t)n_tiles + 1)));
    GNNB_CUDA(cudaMemsetAsync(plan->flag.ptr, 0, sizeof(int), s));
    tc_tile_bounds_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, s>>>(node_ptr, n_graphs, window,
                                                                    n_tiles,
                                                                    plan->bounds.as<int32_t>());
    GNNB_CUDA(cudaGetLastError());
    TcParams p = plan->params;
    p.x = x; p.coo = coo; p.node_ptr = node_ptr; p.edge_ptr = edge_ptr; p.n_graphs = n_graphs;
    p.out = out; p.tile_bounds = plan->bounds.as<int32_t>(); p.n_tiles = n_tiles;
    p.error_flag = plan->flag.as<int>();
    p.scratch = plan->scratch.as<float>();
    p.pending = plan->pending.as<float>();
    p.timing = plan->timing.as<unsigned long long>();
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    fused_tc_kernel<<<grid, NTHREADS, plan->smem_bytes, s>>>(p);
    GNNB_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
    return GNNB_OK;
}

int fused_tc_status(gnnb_model *m, int *status)
{
    *status = 0;
    TcPlan *plan = m->fused_tc;
    if (plan == nullptr) return GNNB_OK;
    GNNB_CUDA(cudaMemcpy(status, plan->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    if (plan->timing.ptr != nullptr) {
        unsigned long long t[16];
        GNNB_CUDA(cudaMemcpy(t, plan->timing.ptr, sizeof(t), cudaMemcpyDeviceToHost));
        GNNB_CUDA(cudaMemset(plan->timing.ptr, 0, sizeof(t)));
        const char *names[8] = {"stage", "tables", "aggregate", "epilogue", "pool", "head",
                                "agg-barrier", "mma"};
        unsigned long long tot = 0;
        for (int i = 0; i < 8; i++) tot += t[i];
        fprintf(stderr, "[gnnb fused-tc phases]");
        for (int i = 0; i < 8; i++)
            fprintf(stderr, " %s %.1f%%", names[i], tot ? 100.0 * (double)t[i] / (double)tot : 0.0);
        fprintf(stderr, " (total %.3g cycles over all CTAs)\n", (double)tot);
    }
    return GNNB_OK;
}

}  // namespace gnnb
