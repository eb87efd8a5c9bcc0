// Kernel (1), tensor-core version: the persistent whole-model fused kernel for molecular-sized
// graphs with BOTH halves of every conv layer on the 5th-generation tensor cores (tcgen05):
//
//   aggregation     AGG = ADJ . X       kind::f16 MMAs.  ADJ is the tile's 128x128 block-diagonal
//                                       matrix of in-edge multiplicities (bf16, exact), X the layer
//                                       input as two bf16 planes (hi + mid, both rounded to
//                                       nearest), fp32 accumulation in tensor memory.
//   node transform  D = A . W^T         "bf16x2": every fp32 operand v ~= hi + mid (two bf16 values,
//                                       |v - hi - mid| <= 2^-18 |v|) and a product is
//                                       A_hi.W_hi + A_mid.W_hi + A_hi.W_mid as three kind::f16 MMAs
//                                       of K = 16.  The A operand is read from TENSOR MEMORY as PACKED
//                                       bf16 pairs (K elements 2j, 2j+1 in the low / high half of
//                                       column j: layout pinned by tools/tmem_bf16_probe.py,
//                                       profiles/r2_tmem_bf16_layout_probe.txt); the weight atoms
//                                       stream L2 -> shared memory with cp.async.bulk.
// The north star asks for <= 1e-4 relative: round 1's 3xTF32 (three kind::tf32 MMAs of K = 8 per
// product, ~2^-21) measured 5e-6 and spent twice the tensor-pipe time and weight bytes; bf16x2
// measures 9.6e-6 worst over all configs and variants (tools/fused_error.py).  -DGNNB_TC_BF2=0
// rebuilds the 3xTF32 kernel (three planes, A_hi / A_lo as fp32 cells) for A/B runs; the comments
// below describe the default.
//
// Every other step is "thread per row": the thread that owns TMEM lane r (row r of the tile) moves
// its row accumulator -> registers (tcgen05.ld) -> bias / skip / activation / GCN-SAGE scaling ->
// either back to tensor memory as the next A operand (tcgen05.st) or to the bf16 planes of the next
// layer input (16-byte shared-memory stores).  GCN / GIN / SAGE: no gather loops, no neighbor
// tables, no shared-memory traffic for activations.
//
// Per CTA (one per SM; 8 worker warps + a weight-producer warp + an MMA-issuing warp; a tile =
// up to 128 rows of whole graphs):
//   shared memory   ADJ  32 KB   bf16 [dst][src], K-major SWIZZLE_128B (A operand of the aggregation)
//                                (PNA: 50 KB of fp32 A_u rows instead)
//                   XP   64 KB   two bf16 planes [node][feature], MN-major SWIZZLE_128B (B operand)
//                   RING 64 KB   4 slots of 16 KB for weight atoms (PNA with N <= 80: 8 x 10 KB)
//                   CNT  16 KB   u8 edge multiplicities built with shared-memory atomics
//   tensor memory   512 columns: D0 [0,128), D1 [128,256), A_hi [384,448), A_mid [448,512);
//                   a PNA layer lays D_id | D_amp | D_att | B_v out in [0,384) itself
// The producer warp issues the bulk copies, the issuing warp all MMAs (warp-uniform code, elected
// lane); completion flows through mbarriers (tcgen05.commit).  The weight stream is continuous
// across GEMMs, layers, the head and tiles.
//
// GCN   planes hold dinv (.) X, ADJ gets +I, the row result is scaled by dinv_v: lib:1246-1278
// GIN   ADJ gets +I, eps * x_v is added in registers: lib:1519-1529; hidden layer stays in TMEM
// SAGE  mean = (ADJ . X) / deg, two transforms accumulate in the same TMEM tile: lib:2180-2207
// PNA   factorised pre-transform, max / min / mean / std gathered thread-per-row from shared
//       memory, the three degree scalers as per-row scalars in front of three accumulators
//       (pna_layer_* below): lib:1750-2157
// Pooling reads the last layer's fp32 rows (warp per graph); the MLP head runs on the tensor cores
// for 128 pooled graphs at a time (pending buffer in L2), cpp:454-530.
//
// Non-finite activations would leak between the graphs of a tile through 0 * Inf inside the
// aggregation MMA (the reference keeps graphs independent), so every value written to the planes
// is checked and the batch is re-run on the layerwise path if any is Inf/NaN (status 3).  PNA has
// no aggregation MMA: its NaN rows (in-degree 0, lib:702) stay in their graph.
//
// Supported: GCN, GIN and SAGE with layer widths a multiple of 16 up to 128, PNA up to 96 (and
// F_in <= F_out); everything else uses fused.cu / the layerwise path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "model.h"
#include "tc.cuh"

namespace gnnb {

// Weight image for the tensor-core path: for every K atom (32 columns of K) the N x 128-byte
// rows of W in the canonical swizzled layout, hi part followed by lo part.
// Rows >= n_valid of the image are zero (N padded up to the MMA granule).
void build_weight_image(const float *W, int N, int n_valid, int K, int ld, int col0,
                        std::vector<float> &img)
{
    const int KA = (K + tc::ATOM_K - 1) / tc::ATOM_K;
    const size_t atom_floats = (size_t)2 * N * tc::ATOM_K;
    const size_t base = img.size();
    img.resize(base + (size_t)KA * atom_floats, 0.0f);
    for (int ka = 0; ka < KA; ka++) {
        float *hi = img.data() + base + (size_t)ka * atom_floats;
        float *lo = hi + (size_t)N * tc::ATOM_K;
        for (int n = 0; n < N; n++)
            for (int kk = 0; kk < tc::ATOM_K; kk++) {
                const int k = ka * tc::ATOM_K + kk;
                const float v = (k < K && n < n_valid) ? W[(size_t)n * ld + col0 + k] : 0.0f;
                const float h = tc::tf32_hi(v);
                const uint32_t off = tc::canon_offset(n, kk, N) / 4;  // atom-local (k < 32)
                hi[off] = h;
                lo[off] = v - h;
            }
    }
}


// bf16 rounded to nearest (ties away from zero in magnitude), as an fp32 bit pattern
static inline uint32_t bf16_rn_bits_host(float v)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc00000u;   // NaN stays NaN
    return (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;      // round to nearest even
}
// bf16x2 weight image: per K atom (64 columns of K) the N x 128-byte rows of W as bf16 in the
// canonical K-major SWIZZLE_128B layout, hi part (W rounded to bf16) followed by the mid part
// (the residual rounded to bf16).  Same byte size per atom as the TF32 image of 32 columns.
void build_weight_image_bf2(const float *W, int N, int n_valid, int K, int ld, int col0,
                            std::vector<float> &img)
{
    const int KA = (K + 63) / 64;
    const size_t atom_floats = (size_t)2 * N * 32;   // 2 units of N x 128 bytes
    const size_t base = img.size();
    img.resize(base + (size_t)KA * atom_floats, 0.0f);
    for (int ka = 0; ka < KA; ka++) {
        uint16_t *hi = reinterpret_cast<uint16_t *>(img.data() + base + (size_t)ka * atom_floats);
        uint16_t *mid = hi + (size_t)N * 64;
        for (int n = 0; n < N; n++)
            for (int kk = 0; kk < 64; kk++) {
                const int k = ka * 64 + kk;
                const float v = (k < K && n < n_valid) ? W[(size_t)n * ld + col0 + k] : 0.0f;
                const uint32_t hb = bf16_rn_bits_host(v);
                float hf;
                memcpy(&hf, &hb, 4);
                const uint32_t mb = bf16_rn_bits_host(v - hf);
                const size_t off = (size_t)n * 64 + (size_t)(((kk >> 3) ^ (n & 7)) * 8) + (kk & 7);
                hi[off] = (uint16_t)(hb >> 16);
                mid[off] = (uint16_t)(mb >> 16);
            }
    }
}

namespace {

constexpr int TM = 128;
#ifndef GNNB_TC_WORKERS
#define GNNB_TC_WORKERS 256
#endif
constexpr int NTHREADS = GNNB_TC_WORKERS;   // worker threads (8 or 16 warps)
constexpr int NWARPS = NTHREADS / 32;
constexpr int CTA_THREADS = NTHREADS + 64;   // + the weight-producer warp + the MMA-issuing warp
constexpr int CSTRIDE = 32 * (NWARPS / 4);   // column stride between the 32-column blocks of one warp
constexpr int NCHALF = NTHREADS / TM;        // staging: threads per row
constexpr int MAX_LAYERS = 8;
constexpr int MAX_HEAD = 6;
constexpr int HEAD_G = 128;      // pooled graphs per head call (one 128-row MMA tile)
constexpr int MAX_HCHUNK = 4;    // 128-wide K chunks of the first head layer (head_in <= 512)
constexpr int MAX_DIM = 128;
constexpr int PLD = 520;         // row stride of the pending pooled vectors (>= 512 + 4)
// Weight ring: 2^slot_log2 slots of slot_stride bytes (a unit = N x 128 bytes).  4 x 16 KB for
// N <= 128; PNA (N <= 80, ~130 units per tile) runs 8 x 10 KB so that more bulk copies are in
// flight: its per-tile weight stream is 1.3 MB and latency-, not bandwidth-, bound.
constexpr int MAX_NSLOT = 8;
constexpr int RING_BYTES_DEFAULT = 4 * 16384;
constexpr int CNT_BYTES = TM * TM;
constexpr int MAX_NODES_PER_GRAPH = TM;   // a graph must fit one tile
#ifndef GNNB_TC_BF2
#define GNNB_TC_BF2 1
#endif
constexpr bool BF2 = GNNB_TC_BF2 != 0;
constexpr int NPLANES = BF2 ? 2 : 3;            // bf16 planes of the layer input
constexpr int WATOM_K = BF2 ? 64 : 32;          // K elements per 128-byte row of a weight atom
constexpr int WMMA_K = BF2 ? 16 : 8;            // K per MMA of the node transform
constexpr uint32_t TMEM_COLS = 512;
// A operand: bf16x2 = packed pairs, 64 columns per 128 K for each of hi / mid; 3xTF32 = 128 each
//   bf16x2: D0 [0,128) D1 [128,256) A_hi [384,448) A_mid [448,512); PNA lays its accumulators out
//           in [0,384) itself (pna_layer_workers)
//   3xTF32: D0 [0,128) A_hi [128,256) A_lo [256,384) D1 [384,512)
constexpr uint32_t TM_D0 = 0, TM_AHI = BF2 ? 384 : 128, TM_ALO = BF2 ? 448 : 256, TM_D1 = BF2 ? 128 : 384;
static_assert(!BF2 || GNNB_TC_WORKERS == 256, "bf16x2 packs 32-column chunks into 16 cells: 8 worker warps");
static_assert(NWARPS == 8 || NWARPS == 16, "row passes: 8 worker warps (two 32-column blocks each) or 16 (one each)");
constexpr int PASS_STEPS = 2;                // one step per 64-column half
constexpr int STEP_COLS = 256 / NWARPS;      // columns of its rows a warp converts per step (32 or 16)
#ifndef GNNB_TC_CHUNK
#define GNNB_TC_CHUNK (GNNB_TC_WORKERS == 512 ? 16 : 32)
#endif
constexpr int CH = GNNB_TC_CHUNK;            // accumulator columns a thread holds at a time (16 keeps 16 warps under 96 registers)
static_assert((CH == 16 || CH == 32) && CH <= STEP_COLS, "row passes work on 16- or 32-column chunks");
// Hand-off arrivals: one per worker WARP and 64-column half (lane 0, after the warp has converged).
// -DGNNB_TC_THREAD_ARRIVALS=1 makes every worker THREAD arrive itself instead -- slower (several
// hundred cycles per phase on one mbarrier word), but an ordering compute-sanitizer's racecheck can
// follow: the debug build for tools/r2 racecheck runs (profiles/r2_sanitizer.txt).
#ifndef GNNB_TC_THREAD_ARRIVALS
#define GNNB_TC_THREAD_ARRIVALS 0
#endif
constexpr int READY_ARRIVALS = GNNB_TC_THREAD_ARRIVALS ? NTHREADS : NWARPS;
// The accumulator is double buffered: MMA phase i writes D[i & 1] (an accumulating phase stays on
// the buffer of the phase it adds to), so the row pass that reads phase i's result can overlap the
// first MMAs of phase i + 1.
__device__ __forceinline__ uint32_t dcol_of(uint32_t dw) { return dw ? TM_D1 : TM_D0; }

struct TLinear {
    const float *img;   // weight image: per K atom [hi N x 128 B | lo N x 128 B]
    const float *bias;
    int K, N, KA;
};

// PNA conv layer (lib:1750-2157) on the tile, factorised as in model.cu (W_pre = [W_self | W_nbr],
// the three degree scalers as per-row scalars in front of three accumulators):
//   A_u = X.W_nbr^T, B_v = X.W_self^T + b_pre, S = X.W_post[:, :F]^T            (three GEMMs, one A operand)
//   stats(v) = max / min / mean / std over in-neighbors u of (A_u + B_v)        (thread per row, shared memory)
//   D_id += stats.W_id^T, D_amp = stats.W_amp^T, D_att = stats.W_att^T          (per group of 32 features)
//   y = lin(S + D_id + amp_v D_amp + att_v D_att + b_post)
constexpr int PNA_LDA = 100;    // row stride (floats) of the A_u rows in shared memory (F_in <= 96, + 4)
constexpr int MAX_PNA_LAYERS = 4;
constexpr int MAX_PNA_GROUPS = 3;     // feature groups of 32: F_in <= 96
struct PnaLayer {
    TLinear pa, pb, ps;                        // X -> A_u, X -> B_v (bias b_pre), X -> S
    TLinear gid[MAX_PNA_GROUPS], gamp[MAX_PNA_GROUPS], gatt[MAX_PNA_GROUPS];
    TLinear pl;                                // final linear (bias b_lin)
    const float *b_post;
    int ng, gw[MAX_PNA_GROUPS], fiP, foP;
};

struct TcParams {
    int conv_type, num_layers, in_dim, skip, gnn_act, num_pools, pools[4];
    int mlp_num_linear, mlp_act, out_act, emb, mlp_out;
    float gin_eps;
    int fi[MAX_LAYERS], fo[MAX_LAYERS];
    TLinear l0[MAX_LAYERS], l1[MAX_LAYERS];   // l1: GIN second linear / SAGE root weight
    TLinear hl[MAX_HEAD][MAX_HCHUNK];         // MLP head linears, K cut into 128-wide chunks
    int hchunks[MAX_HEAD];
    int head_n[MAX_HEAD];                     // true output widths (N of the images is padded to 16)
    const float *x;
    const int32_t *coo;
    const int64_t *node_ptr, *edge_ptr;
    int n_graphs;
    float *out;
    const int32_t *tile_bounds;   // [n_tiles + 1] first graph of every tile (device-built)
    const int32_t *n_tiles_ptr;   // number of tiles (device-built)
    int *error_flag;
    float *pending;               // [grid][HEAD_G][PLD] pooled vectors waiting for the head
    unsigned long long *timing;
    size_t img_copy_bytes;        // distance between the replicas of the weight images
    int img_copies;
    int r0_bytes;                 // first shared-memory region: ADJ (32 KB) or PNA's A_u rows
    int ring_bytes, slot_log2, slot_stride;   // weight ring geometry
    float pna_delta;
    PnaLayer pna[MAX_PNA_LAYERS];
};

struct Misc {
    uint64_t bar_full[MAX_NSLOT], bar_empty[MAX_NSLOT];
    uint32_t slot_log2, slot_stride;
    unsigned long long *timing;      // GNNB_FUSED_TIMING: per-phase cycle counters, else null
    // MMA <-> row-pass hand-off, per 64-column half h: ready[h] = the workers have written columns
    // [64 h, 64 h + 64) of the next MMA operand (one arrival per worker warp); done[h] = the MMAs
    // producing columns [64 h, +64) of the accumulator have completed (tcgen05.commit).
    uint64_t bar_done[2], bar_ready[2];
    uint32_t tmem_slot;
    int nonfinite;
    int pend_gid[HEAD_G];
    int deg[TM];
    int grow[TM + 2];
    int gedge[TM + 2];
};

// Tile packing on the device: consecutive graphs are packed greedily into tiles of <= 128 rows
// (and <= 128 graphs).  The greedy walk is sequential, so the batch is cut into chunks of
// PACK_CHUNK graphs that are packed independently (one CTA each: every thread finds, for "its"
// graph, where a tile starting there would end; one thread then follows those links), followed
// by a scan of the per-chunk tile counts and a compaction into the final tile list.  A chunk
// boundary costs at most one partly filled tile per 2048 graphs.
constexpr int PACK_CHUNK = 2048;
__global__ void __launch_bounds__(256) tc_pack_chunk_kernel(const int64_t *__restrict__ node_ptr,
                                                            int n_graphs, int32_t *__restrict__ starts_tmp,
                                                            int32_t *__restrict__ counts)
{
    __shared__ int np[PACK_CHUNK + 1];
    __shared__ unsigned short nxt[PACK_CHUNK];
    const int g0 = blockIdx.x * PACK_CHUNK;
    const int cnt = min(PACK_CHUNK, n_graphs - g0);
    const int64_t base = __ldg(node_ptr + g0);
    for (int i = threadIdx.x; i <= cnt; i += blockDim.x) {
        const int64_t d = __ldg(node_ptr + g0 + i) - base;
        np[i] = d > (int64_t)0x3fffffff ? 0x3fffffff : (int)d;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        // largest e in [i + 1, min(cnt, i + TM)] with np[e] - np[i] <= TM (i + 1 always taken)
        int lo = i + 1, hi = min(cnt, i + TM);
        const int lim = np[i] + TM;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (np[mid] <= lim) lo = mid; else hi = mid - 1;
        }
        nxt[i] = (unsigned short)lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int pos = 0; pos < cnt; pos = nxt[pos]) starts_tmp[(size_t)blockIdx.x * PACK_CHUNK + n++] = g0 + pos;
        counts[blockIdx.x] = n;
    }
}
// exclusive scan of the per-chunk tile counts (one CTA), total -> *n_tiles
__global__ void __launch_bounds__(1024) tc_pack_scan_kernel(const int32_t *__restrict__ counts,
                                                            int n_chunks, int32_t *__restrict__ offsets,
                                                            int32_t *__restrict__ n_tiles)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < n_chunks; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int v = i < n_chunks ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - v;
        if (i < n_chunks) offsets[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_tiles = carry;
}
__global__ void __launch_bounds__(256) tc_pack_compact_kernel(const int32_t *__restrict__ starts_tmp,
                                                              const int32_t *__restrict__ counts,
                                                              const int32_t *__restrict__ offsets,
                                                              int n_chunks, int n_graphs,
                                                              int32_t *__restrict__ bounds)
{
    const int c = blockIdx.x;
    const int n = counts[c], off = offsets[c];
    for (int i = threadIdx.x; i < n; i += blockDim.x) bounds[off + i] = starts_tmp[(size_t)c * PACK_CHUNK + i];
    if (c == n_chunks - 1 && threadIdx.x == 0) bounds[off + n] = n_graphs;
}

// ---------------------------------------------------------------------------------------
// Weight stream.  A dedicated producer warp walks the (static) sequence of linears the CTA will
// run -- per tile: the conv layers' linears in order, then the head's if 128 pooled graphs are
// waiting -- and keeps the 4-slot ring full of weight "units" (one 32-wide K atom of the hi or
// the lo image, N x 128 bytes) with TMA bulk copies; the MMA-issuing warp consumes units in the
// same order.  full[s] / empty[s] mbarriers carry the hand-off, so the stream runs ahead across
// GEMM, layer and tile boundaries without any bookkeeping on the consumer side.
//
// Both warps execute warp-uniform control flow (every value they use is broadcast from lane 0 or
// comes from the kernel parameters), so the compiler keeps descriptors and addresses in uniform
// registers and a tcgen05.mma costs one instruction; only the MMA / commit / bulk-copy
// instructions themselves are predicated on the elected lane.  Issued from a divergent
// `if (tid == 0)` region the same loop costs ~110 cycles per MMA (a register -> uniform-register
// waterfall per instruction), 1.7x the MMA itself (tools/mma_rate.py).
// Unit order of one linear: the hi atoms 0 .. KA-1, then the lo atoms.  (Measured: moving the lo
// atoms of the first two K atoms forward, so that half of the GEMM could start on the first half of
// the A operand, is 3 % slower -- with a 4-slot ring the second group's units could then only be
// requested once the first group's MMAs had completed.)
__device__ __forceinline__ void produce_linear(Misc &ms, uint32_t ring, uint32_t &prod,
                                               const TLinear &L, size_t copy_off, bool leader)
{
    const uint32_t bytes = (uint32_t)L.N * tc::ROW_BYTES;
    const uint32_t slot_log2 = ms.slot_log2, slot_stride = ms.slot_stride;
    const unsigned char *hi = reinterpret_cast<const unsigned char *>(L.img) + copy_off;
    const unsigned char *lo = hi + bytes;
    for (int part = 0; part < 2; part++) {
        const unsigned char *src = part ? lo : hi;
        for (int ka = 0; ka < L.KA; ka++) {
            const uint32_t s = prod & ((1u << slot_log2) - 1u), use = prod >> slot_log2;
            if (use > 0) tc::mbar_wait(&ms.bar_empty[s], (use - 1) & 1);
            if (leader) {
                tc::mbar_expect_tx(&ms.bar_full[s], bytes);
                tc::bulk_g2s_addr(ring + s * slot_stride, src, bytes, &ms.bar_full[s]);
            }
            src += 2 * (size_t)bytes;
            prod++;
        }
    }
}
// D[128][N] (+)= A(hi,lo in tensor memory)[128][KA*32] . W^T, 3xTF32: first the hi weight atoms
// (A_hi.B_hi and A_lo.B_hi), then the lo atoms (A_hi.B_lo).  All lanes of the issuing warp.
// The A operand arrives in two halves: the hi atoms that only touch columns [0, 64) are issued as
// soon as ready[0] completes, i.e. while the workers are still converting columns [64, 128).
#if GNNB_TC_BF2
// bf16x2 version: a weight atom covers K = 64 (= one 64-column half of the A operand, i.e. 32
// packed TMEM columns per part), 4 k-steps of K = 16; hi atoms feed A_hi.W_hi and A_mid.W_hi, the
// mid atoms A_hi.W_mid.  Only the k-steps that cover real K columns are issued (layer 0 of the
// QM9 model has K = 11: one k-step).
// wait_ready = false: the A operand is the one the previous GEMM used (several GEMMs off one row
// pass); commit_done = false: more GEMMs of the same phase follow before the workers are told.
__device__ __forceinline__ void gemm_issue(Misc &ms, uint32_t ring, uint32_t &cons, uint32_t tmem_base,
                                           uint32_t dcol, const TLinear &L, bool accumulate,
                                           uint32_t &ready_cnt, bool wait_ready = true,
                                           bool commit_done = true)
{
    const bool leader = tc::elect_one();
    const uint32_t slot_log2 = ms.slot_log2, slot_stride = ms.slot_stride;
    const int KA = L.KA;
    const uint32_t idesc = tc::make_idesc_bf16(TM, L.N, 0);
    const uint32_t tmem_d = tmem_base + dcol, ahi = tmem_base + TM_AHI, amid = tmem_base + TM_ALO;
    // GNNB_FUSED_TIMING: the issuing warp's cycles waiting for the A operand (timing[16]), waiting
    // for weight units (timing[17]) and inside this function in total (timing[18])
    unsigned long long *const timing = ms.timing;
    long long t_enter = 0, t_w = 0, t_ready = 0, t_full = 0;
    if (timing != nullptr) t_enter = clock64();
#define GNNB_TWAIT(acc, expr) do { if (timing != nullptr) t_w = clock64(); expr; if (timing != nullptr) acc += clock64() - t_w; } while (0)
    if (wait_ready) {
        GNNB_TWAIT(t_ready, tc::mbar_wait(&ms.bar_ready[0], ready_cnt & 1));
        tc::tc_fence_after();
    }
    for (int ka = 0; ka < KA; ka++) {   // hi atoms
        if (ka == 1 && wait_ready) {
            GNNB_TWAIT(t_ready, tc::mbar_wait(&ms.bar_ready[1], ready_cnt & 1));
            tc::tc_fence_after();
        }
        const uint32_t s = cons & ((1u << slot_log2) - 1u);
        GNNB_TWAIT(t_full, tc::mbar_wait(&ms.bar_full[s], (cons >> slot_log2) & 1));
        tc::tc_fence_after();
        const uint32_t col = (uint32_t)(ka * 32);
        const int nk = min(4, (L.K - ka * 64 + 15) >> 4);
        const uint64_t bd = tc::make_desc(ring + s * slot_stride);
        const uint32_t first = (ka == 0 && !accumulate) ? 0u : 1u;
        if (leader) {
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                if (k4 < nk) {
                    // +2 per k-step: 16 bf16 = 32 bytes along K inside the swizzle atom (>> 4)
                    tc::mma_bf16_ts(tmem_d, ahi + col + 8 * k4, bd + (uint64_t)(2 * k4), idesc,
                                    k4 == 0 ? first : 1u);
                    tc::mma_bf16_ts(tmem_d, amid + col + 8 * k4, bd + (uint64_t)(2 * k4), idesc, 1u);
                }
            }
            tc::mma_commit(&ms.bar_empty[s]);
        }
        cons++;
    }
    if (wait_ready) {
        if (KA <= 1) {
            GNNB_TWAIT(t_ready, tc::mbar_wait(&ms.bar_ready[1], ready_cnt & 1));
            tc::tc_fence_after();
        }
        ready_cnt++;
    }
    for (int ka = 0; ka < KA; ka++) {   // mid atoms
        const uint32_t s = cons & ((1u << slot_log2) - 1u);
        GNNB_TWAIT(t_full, tc::mbar_wait(&ms.bar_full[s], (cons >> slot_log2) & 1));
        tc::tc_fence_after();
        const uint32_t col = (uint32_t)(ka * 32);
        const int nk = min(4, (L.K - ka * 64 + 15) >> 4);
        const uint64_t bd = tc::make_desc(ring + s * slot_stride);
        if (leader) {
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++)
                if (k4 < nk)
                    tc::mma_bf16_ts(tmem_d, ahi + col + 8 * k4, bd + (uint64_t)(2 * k4), idesc, 1u);
            tc::mma_commit(&ms.bar_empty[s]);
        }
        cons++;
    }
    if (leader && commit_done) {
        tc::mma_commit(&ms.bar_done[0]);
        tc::mma_commit(&ms.bar_done[1]);
    }
#undef GNNB_TWAIT
    if (timing != nullptr && leader) {
        atomicAdd(timing + 16, (unsigned long long)t_ready);
        atomicAdd(timing + 17, (unsigned long long)t_full);
        atomicAdd(timing + 18, (unsigned long long)(clock64() - t_enter));
        atomicAdd(timing + 19, 1ull);
    }
}
#else
__device__ __forceinline__ void gemm_issue(Misc &ms, uint32_t ring, uint32_t &cons, uint32_t tmem_base,
                                           uint32_t dcol, const TLinear &L, bool accumulate,
                                           uint32_t &ready_cnt)
{
    const bool leader = tc::elect_one();
    const uint32_t slot_log2 = ms.slot_log2, slot_stride = ms.slot_stride;
    const int KA = L.KA;
    const int KA0 = KA < 2 ? KA : 2;   // atoms inside the first 64 columns
    const uint32_t idesc = tc::make_idesc_tf32(TM, L.N);
    const uint32_t tmem_d = tmem_base + dcol, ahi = tmem_base + TM_AHI, alo = tmem_base + TM_ALO;
    tc::mbar_wait(&ms.bar_ready[0], ready_cnt & 1);
    tc::tc_fence_after();
    for (int ka = 0; ka < KA; ka++) {   // hi atoms
        if (ka == KA0) {
            tc::mbar_wait(&ms.bar_ready[1], ready_cnt & 1);
            tc::tc_fence_after();
        }
        const uint32_t s = cons & ((1u << slot_log2) - 1u);
        tc::mbar_wait(&ms.bar_full[s], (cons >> slot_log2) & 1);
        tc::tc_fence_after();
        const uint32_t col = (uint32_t)(ka * tc::ATOM_K);
        const uint64_t bd = tc::make_desc(ring + s * slot_stride);
        const uint32_t first = (ka == 0 && !accumulate) ? 0u : 1u;
        if (leader) {
#pragma unroll
            for (int k8 = 0; k8 < tc::ATOM_K / tc::MMA_K; k8++) {
                // +2 per k-step: 8 TF32 = 32 bytes along K inside the swizzle atom (>> 4)
                tc::mma_tf32_ts(tmem_d, ahi + col + k8 * tc::MMA_K, bd + (uint64_t)(2 * k8), idesc,
                                k8 == 0 ? first : 1u);
                tc::mma_tf32_ts(tmem_d, alo + col + k8 * tc::MMA_K, bd + (uint64_t)(2 * k8), idesc, 1u);
            }
            tc::mma_commit(&ms.bar_empty[s]);
        }
        cons++;
    }
    if (KA <= KA0) {
        tc::mbar_wait(&ms.bar_ready[1], ready_cnt & 1);
        tc::tc_fence_after();
    }
    ready_cnt++;
    for (int ka = 0; ka < KA; ka++) {   // lo atoms
        const uint32_t s = cons & ((1u << slot_log2) - 1u);
        tc::mbar_wait(&ms.bar_full[s], (cons >> slot_log2) & 1);
        tc::tc_fence_after();
        const uint32_t col = (uint32_t)(ka * tc::ATOM_K);
        const uint64_t bd = tc::make_desc(ring + s * slot_stride);
        if (leader) {
#pragma unroll
            for (int k8 = 0; k8 < tc::ATOM_K / tc::MMA_K; k8++)
                tc::mma_tf32_ts(tmem_d, ahi + col + k8 * tc::MMA_K, bd + (uint64_t)(2 * k8), idesc, 1u);
            tc::mma_commit(&ms.bar_empty[s]);
        }
        cons++;
    }
    if (leader) {
        tc::mma_commit(&ms.bar_done[0]);
        tc::mma_commit(&ms.bar_done[1]);
    }
}
#endif

// AGG[128][n_cols] = ADJ . (Xh + Xm + Xl), K = source nodes (16 per MMA, only the k-steps that
// cover the tile's rows).  Warp-uniform, like gemm_issue.  Issued per 64-column half of the
// features (one 64-feature block of the planes each): half 0 starts when the workers have written
// plane columns [0, 64) and is handed back (done[0]) while half 1 is still running.
__device__ __forceinline__ void agg_issue(Misc &ms, uint32_t tmem_d, uint32_t adj, uint32_t xp,
                                          int n_cols, int rows, uint32_t &ready_cnt)
{
    const bool leader = tc::elect_one();
    const int nks = rows > 0 ? (rows + 15) >> 4 : 1;   // (an all-empty tile still defines D)
    // ADJ: k-step ks lives in K atom ks >> 2 at byte 32 (ks & 3); planes: 16 nodes = 2048 bytes
    const uint64_t a0 = tc::make_desc(adj), a1 = tc::make_desc(adj + tc::PLANE_BLOCK_BYTES);
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        tc::mbar_wait(&ms.bar_ready[h], ready_cnt & 1);
        tc::tc_fence_after();
        const int nh = min(n_cols - 64 * h, 64);
        if (nh > 0) {
            const uint32_t idesc = tc::make_idesc_bf16(TM, nh, 1);
#pragma unroll
            for (int pl = 0; pl < NPLANES; pl++) {
                const uint64_t b0 = tc::make_desc_mn(xp + (uint32_t)pl * tc::PLANE_BYTES +
                                                         (uint32_t)h * tc::PLANE_BLOCK_BYTES,
                                                     tc::PLANE_BLOCK_BYTES, 1024u);
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    if (ks < nks && leader)
                        tc::mma_bf16(tmem_d + 64u * (uint32_t)h, (ks < 4 ? a0 : a1) + (uint64_t)(2 * (ks & 3)),
                                     b0 + (uint64_t)(128 * ks), idesc, (pl == 0 && ks == 0) ? 0u : 1u);
                }
            }
        }
        if (leader) tc::mma_commit(&ms.bar_done[h]);
    }
    ready_cnt++;
}

// ---------------------------------------------------------------------------------------
// Thread-per-row helpers.  Warp w owns TMEM lanes [32 (w & 3), +32) and the column blocks
// c0 = 32 (w >> 2), +64, ...
// two fp32 values -> one packed bf16 pair (lo in the low half), round to nearest even, NaN kept:
// one cvt.rn.bf16x2.f32.  (An integer add-and-mask rounding turns CUDA's canonical NaN 0x7fffffff
// into -0.0: a zero-in-degree PNA row must stay NaN like the reference's, lib:702.)
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// hi / mid bf16 pairs of two fp32 values: v ~= hi + mid, both rounded to nearest
__device__ __forceinline__ void split2_pair(float v0, float v1, uint32_t &hp, uint32_t &mp)
{
    hp = pack_bf16x2_rn(v0, v1);
    const float r0 = v0 - __uint_as_float(hp << 16), r1 = v1 - __uint_as_float(hp & 0xffff0000u);
    mp = pack_bf16x2_rn(r0, r1);
}
__device__ __forceinline__ void load_row8(const unsigned char *XP, int row, int c, float (&v)[8])
{
    const uint32_t off = tc::plane_chunk_offset(row, c);
    const uint4 h = *reinterpret_cast<const uint4 *>(XP + off);
    const uint4 m = *reinterpret_cast<const uint4 *>(XP + tc::PLANE_BYTES + off);
    if (BF2) {
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(mw[j] << 16);
            v[2 * j + 1] = __uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(mw[j] & 0xffff0000u);
        }
    } else {
        const uint4 l = *reinterpret_cast<const uint4 *>(XP + 2 * tc::PLANE_BYTES + off);
        tc::join3_unpack8(h, m, l, v);
    }
}
__device__ __forceinline__ void store_row8(unsigned char *XP, int row, int c, const float (&v)[8])
{
    const uint32_t off = tc::plane_chunk_offset(row, c);
    if (BF2) {
        uint32_t hp[4], mp[4];
#pragma unroll
        for (int j = 0; j < 4; j++) split2_pair(v[2 * j], v[2 * j + 1], hp[j], mp[j]);
        *reinterpret_cast<uint4 *>(XP + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4 *>(XP + tc::PLANE_BYTES + off) = make_uint4(mp[0], mp[1], mp[2], mp[3]);
    } else {
        uint4 h, m, l;
        tc::split3_pack8(v, h, m, l);
        *reinterpret_cast<uint4 *>(XP + off) = h;
        *reinterpret_cast<uint4 *>(XP + tc::PLANE_BYTES + off) = m;
        *reinterpret_cast<uint4 *>(XP + 2 * tc::PLANE_BYTES + off) = l;
    }
}
// W consecutive fp32 columns [c0, c0 + W) of this thread's row -> the A operand in tensor memory.
// 3xTF32: hi / lo parts, one 32-bit cell per element.  bf16x2: hi / mid parts as packed bf16
// pairs (elements 2j, 2j+1 of the row in the low / high half of cell j), so W columns are W / 2 cells.
template <int W>
__device__ __forceinline__ void split_store(uint32_t tmem_base, uint32_t lane_base, int c0, const float (&v)[W])
{
    if (BF2) {
        float h[W / 2], m[W / 2];
#pragma unroll
        for (int j = 0; j < W / 2; j++) {
            uint32_t hp, mp;
            split2_pair(v[2 * j], v[2 * j + 1], hp, mp);
            h[j] = __uint_as_float(hp);
            m[j] = __uint_as_float(mp);
        }
        tc::tmem_st(tmem_base + TM_AHI + lane_base + (uint32_t)(c0 >> 1), h);
        tc::tmem_st(tmem_base + TM_ALO + lane_base + (uint32_t)(c0 >> 1), m);
    } else {
        float h[W], l[W];
#pragma unroll
        for (int j = 0; j < W; j++) { h[j] = tc::tf32_hi(v[j]); l[j] = v[j] - h[j]; }
        tc::tmem_st(tmem_base + TM_AHI + lane_base + (uint32_t)c0, h);
        tc::tmem_st(tmem_base + TM_ALO + lane_base + (uint32_t)c0, l);
    }
}

// Hand-off to the issuing warp, one arrival per worker warp and per 64-column half: every lane has
// made its part of the operands visible (tcgen05.wait::st, or fence.proxy.async for shared memory),
// the warp converges and lane 0 arrives.  256 single-thread arrivals on one mbarrier word used to
// serialise for several hundred cycles per phase.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void handoff_half(Misc &ms, int h)
{
    tc::tc_fence_before();
#if GNNB_TC_THREAD_ARRIVALS
    mbar_arrive(&ms.bar_ready[h]);
#else
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&ms.bar_ready[h]);
#endif
}
__device__ __forceinline__ void handoff_both(Misc &ms)
{
    tc::tc_fence_before();
#if GNNB_TC_THREAD_ARRIVALS
    mbar_arrive(&ms.bar_ready[0]);
    mbar_arrive(&ms.bar_ready[1]);
#else
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        mbar_arrive(&ms.bar_ready[0]);
        mbar_arrive(&ms.bar_ready[1]);
    }
#endif
}
// wait for both halves of the current MMA phase (every worker polls: no CTA barrier)
__device__ __forceinline__ void wait_done_both(Misc &ms, uint32_t &done_cnt)
{
    tc::mbar_wait(&ms.bar_done[0], done_cnt & 1);
    tc::mbar_wait(&ms.bar_done[1], done_cnt & 1);
    done_cnt++;
    tc::tc_fence_after();
}

// One pass over the accumulator of the current MMA phase, in two steps: step h covers columns
// [64 h, 64 h + 64), STEP_COLS of them per warp (8 worker warps: warps 0-3 the first 32 columns of
// their rows, warps 4-7 the second 32), limited to ncols.  Step h waits for done[h] (the MMAs of that accumulator half),
// runs f(c0, lane_base, v) on the block and, when the pass produces the next MMA operand
// (HK = 1: in tensor memory, HK = 2: in shared memory), hands that half off on ready[h] -- so the
// next phase's first MMAs run while step 1 is still converting, and (aggregation) step 0 runs
// while the second half of the accumulator is still being computed.
#ifdef GNNB_TC_SUBTIMING
__device__ unsigned long long g_sub[8];
#define SUBT(i, expr) do { const long long t0_ = clock64(); expr; if (threadIdx.x == 0) atomicAdd(&g_sub[i], (unsigned long long)(clock64() - t0_)); } while (0)
#else
#define SUBT(i, expr) do { expr; } while (0)
#endif
template <int HK, class F>
__device__ __forceinline__ void row_pass(Misc &ms, uint32_t &done_cnt, uint32_t tmem_src, int ncols, F &&f)
{
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
#pragma unroll 1
    for (int st = 0; st < PASS_STEPS; st++) {
        const int h = st, c0 = 64 * st + STEP_COLS * (warp >> 2);
        SUBT(3, tc::mbar_wait(&ms.bar_done[h], done_cnt & 1));
        tc::tc_fence_after();
        if (c0 < ncols) {
#pragma unroll 1
            for (int cc = c0; cc < c0 + STEP_COLS; cc += CH) {
                uint32_t r[CH];
                SUBT(0, tc::tmem_ld_nowait(tmem_src + lane_base + (uint32_t)cc, r); tc::tmem_ld_wait());
                SUBT(1, f(cc, lane_base, r));
            }
        }
        if (HK == 1) SUBT(2, tc::tmem_st_wait());
        if (HK == 2) tc::fence_async_smem();
        if (HK != 0) handoff_half(ms, h);
    }
    done_cnt++;
}
template <int ACT>   // 0 identity, 1 relu, 2 anything else (one out-of-line call)
__device__ __forceinline__ float act_fast(int act, float x)
{
    if (ACT == 0) return x;
    if (ACT == 1) return fmaxf(x, 0.0f);   // NaN -> 0 like the reference's (x > 0 ? x : 0)
    return act_apply_general(act, x);
}

// aggregation accumulator -> A operand.  MODE 0: v * scale (GCN dinv_v; SAGE 1 / deg, lib:2180-2207),
// MODE 1: v as it is (GIN with eps = 0), MODE 2: v + self_coef * x_v (GIN eps).  Columns [0, kp).
template <int MODE>
__device__ __forceinline__ void cvt_agg(Misc &ms, uint32_t &done_cnt, uint32_t tmem_base, uint32_t dcol,
                                        const unsigned char *XP, int kp, float scale, float self_coef)
{
    const int row = 32 * ((threadIdx.x >> 5) & 3) + (threadIdx.x & 31);
    row_pass<1>(ms, done_cnt, tmem_base + dcol, kp, [&](int c0, uint32_t lane_base, const uint32_t (&r)[CH]) {
        float v[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) v[j] = __uint_as_float(r[j]);
        if (MODE == 2) {
#pragma unroll
            for (int j8 = 0; j8 < CH / 8; j8++) {
                float xs[8];
                load_row8(XP, row, c0 + 8 * j8, xs);
#pragma unroll
                for (int j = 0; j < 8; j++) v[8 * j8 + j] = fmaf(self_coef, xs[j], v[8 * j8 + j]);
            }
        } else if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < CH; j++) v[j] *= scale;
        }
        split_store(tmem_base, lane_base, c0, v);
    });
}

// own row of the planes -> A operand (SAGE root term), columns [0, kp).  The caller has waited for
// the GEMM that was reading the A operand.
__device__ __forceinline__ void cvt_self(Misc &ms, uint32_t tmem_base, const unsigned char *XP, int kp)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = 32 * (warp & 3) + lane;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
#pragma unroll 1
    for (int st = 0; st < PASS_STEPS; st++) {
        const int h = st, c0 = 64 * st + STEP_COLS * (warp >> 2);
        if (c0 < kp) {
#pragma unroll 1
            for (int cc = c0; cc < c0 + STEP_COLS; cc += CH) {
                float v[CH];
#pragma unroll
                for (int j8 = 0; j8 < CH / 8; j8++) {
                    float xs[8];
                    load_row8(XP, row, cc + 8 * j8, xs);
#pragma unroll
                    for (int j = 0; j < 8; j++) v[8 * j8 + j] = xs[j];
                }
                split_store(tmem_base, lane_base, cc, v);
            }
        }
        tc::tmem_st_wait();
        handoff_half(ms, h);
    }
}

// accumulator -> (+bias, activation) -> A operand (GIN hidden layer, head hidden layers).
// Columns [N, round_up(N, 32)) are zero filled.
template <int ACT>
__device__ __forceinline__ void epilogue_tmem_t(Misc &ms, uint32_t &done_cnt, uint32_t tmem_base,
                                                uint32_t dcol, int N, const float *__restrict__ bias, int act)
{
    row_pass<1>(ms, done_cnt, tmem_base + dcol, (N + 31) & ~31, [&](int c0, uint32_t lane_base, const uint32_t (&r)[CH]) {
        float v[CH];
#pragma unroll
        for (int j4 = 0; j4 < CH / 4; j4++) {
            const bool in_range = c0 + j4 * 4 < N;   // N % 4 == 0
            const float4 bs = in_range ? __ldg(reinterpret_cast<const float4 *>(bias + c0 + j4 * 4))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            const float bss[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
            for (int j = 0; j < 4; j++)
                v[j4 * 4 + j] = in_range ? act_fast<ACT>(act, __uint_as_float(r[j4 * 4 + j]) + bss[j]) : 0.0f;
        }
        split_store(tmem_base, lane_base, c0, v);
    });
}
__device__ __forceinline__ void epilogue_tmem(Misc &ms, uint32_t &done_cnt, uint32_t tmem_base, uint32_t dcol,
                                              int N, const float *__restrict__ bias, int act)
{
    if (act == GNNB_ACT_RELU) epilogue_tmem_t<1>(ms, done_cnt, tmem_base, dcol, N, bias, act);
    else if (act == GNNB_ACT_IDENTITY) epilogue_tmem_t<0>(ms, done_cnt, tmem_base, dcol, N, bias, act);
    else epilogue_tmem_t<2>(ms, done_cnt, tmem_base, dcol, N, bias, act);
}

// accumulator -> (+bias, +skip, activation, * out_scale) -> bf16 planes (the next layer's input).
// skip: add the row's current plane content times skip_unscale (cpp:269-279).  Returns nonzero if
// a non-finite value was written (chk accumulates t * 0, which is NaN exactly then).  Columns
// [N, round_up(N, 32)) are zero filled.
template <int ACT, bool SKIP, bool SCALE, int HK = 2>
__device__ __forceinline__ int epilogue_planes_t(Misc &ms, uint32_t &done_cnt, uint32_t tmem_d,
                                                 unsigned char *XP, int N,
                                                 const float *__restrict__ bias, int act,
                                                 float skip_unscale, float out_scale)
{
    const int row = 32 * ((threadIdx.x >> 5) & 3) + (threadIdx.x & 31);
    float chk = 0.0f;
    row_pass<HK>(ms, done_cnt, tmem_d, (N + 31) & ~31, [&](int c0, uint32_t, const uint32_t (&r)[CH]) {
#pragma unroll
        for (int j8 = 0; j8 < CH / 8; j8++) {
            const int c = c0 + 8 * j8;
            float o[8];
            if (c < N) {   // N % 8 == 0
                const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + c));
                const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + c + 4));
                const float bss[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float xs[8];
                if (SKIP) load_row8(XP, row, c, xs);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float t = __uint_as_float(r[8 * j8 + j]) + bss[j];
                    if (SKIP) t = SCALE ? fmaf(xs[j], skip_unscale, t) : t + xs[j];
                    t = act_fast<ACT>(act, t);
                    if (SCALE) t *= out_scale;
                    chk = fmaf(t, 0.0f, chk);
                    o[j] = t;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) o[j] = 0.0f;
            }
            store_row8(XP, row, c, o);
        }
    });
    return chk == 0.0f ? 0 : 1;
}
#if GNNB_TC_BF2
// PNA: the same pass writes the planes (the next layer's skip source) AND the next layer's A operand
// -- the packed bf16 pairs are identical -- and hands the A operand off, so the next layer's three
// pre-transform GEMMs start without another pass over the planes.
template <int ACT, bool SKIP>
__device__ __forceinline__ void epilogue_planes_a_t(Misc &ms, uint32_t &done_cnt, uint32_t tmem_base,
                                                    uint32_t tmem_d, unsigned char *XP, int N,
                                                    const float *__restrict__ bias, int act)
{
    const int row = 32 * ((threadIdx.x >> 5) & 3) + (threadIdx.x & 31);
    row_pass<1>(ms, done_cnt, tmem_d, (N + 31) & ~31, [&](int c0, uint32_t lane_base, const uint32_t (&r)[CH]) {
        float hc[CH / 2], mc[CH / 2];
#pragma unroll
        for (int j8 = 0; j8 < CH / 8; j8++) {
            const int c = c0 + 8 * j8;
            float o[8];
            if (c < N) {   // N % 8 == 0
                const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + c));
                const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + c + 4));
                const float bss[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float xs[8];
                if (SKIP) load_row8(XP, row, c, xs);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float t = __uint_as_float(r[8 * j8 + j]) + bss[j];
                    if (SKIP) t += xs[j];
                    o[j] = act_fast<ACT>(act, t);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) o[j] = 0.0f;
            }
            uint32_t hp[4], mp[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                split2_pair(o[2 * j], o[2 * j + 1], hp[j], mp[j]);
                hc[4 * j8 + j] = __uint_as_float(hp[j]);
                mc[4 * j8 + j] = __uint_as_float(mp[j]);
            }
            const uint32_t off = tc::plane_chunk_offset(row, c);
            *reinterpret_cast<uint4 *>(XP + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            *reinterpret_cast<uint4 *>(XP + tc::PLANE_BYTES + off) = make_uint4(mp[0], mp[1], mp[2], mp[3]);
        }
        tc::tmem_st(tmem_base + TM_AHI + lane_base + (uint32_t)(c0 >> 1), hc);
        tc::tmem_st(tmem_base + TM_ALO + lane_base + (uint32_t)(c0 >> 1), mc);
    });
}
#endif
// SCALE (GCN only): skip_unscale undoes the dinv factor stored in the planes, out_scale applies the
// next layer's.  The common cases (ReLU / identity, with and without skip) get lean instantiations.
__device__ __forceinline__ int epilogue_planes(Misc &ms, uint32_t &done_cnt, uint32_t tmem_d,
                                               unsigned char *XP, int N,
                                               const float *__restrict__ bias, int act, bool skip,
                                               bool scale, float skip_unscale, float out_scale)
{
#define GNNB_EP(A, S, C, u, o) epilogue_planes_t<A, S, C>(ms, done_cnt, tmem_d, XP, N, bias, act, u, o)
    if (scale) {
        if (act == GNNB_ACT_RELU)
            return skip ? GNNB_EP(1, true, true, skip_unscale, out_scale) : GNNB_EP(1, false, true, skip_unscale, out_scale);
        return skip ? GNNB_EP(2, true, true, skip_unscale, out_scale) : GNNB_EP(2, false, true, skip_unscale, out_scale);
    }
    if (act == GNNB_ACT_RELU)
        return skip ? GNNB_EP(1, true, false, 1.0f, 1.0f) : GNNB_EP(1, false, false, 1.0f, 1.0f);
    return skip ? GNNB_EP(2, true, false, 1.0f, 1.0f) : GNNB_EP(2, false, false, 1.0f, 1.0f);
#undef GNNB_EP
}

// Last conv layer: accumulator -> (+bias, +skip, activation) -> plain fp32 rows in the (now dead)
// plane region, for pooling only: row r at 512 r bytes, 16-byte chunk c stored at position
// c ^ (r & 7) so that thread-per-row stores and lane-per-chunk pooling reads are both conflict free.
// No non-finite check: pooling and the head never mix graphs.
__device__ __forceinline__ uint32_t out_chunk_offset(int row, int c4)   // c4 % 4 == 0
{
    return (uint32_t)row * 512u + (uint32_t)(((c4 >> 2) ^ (row & 7)) << 4);
}
template <int ACT>
__device__ __forceinline__ void epilogue_rows_t(Misc &ms, uint32_t &done_cnt, uint32_t tmem_d,
                                                unsigned char *XP, int N,
                                                const float *__restrict__ bias, int act)
{
    const int row = 32 * ((threadIdx.x >> 5) & 3) + (threadIdx.x & 31);
    row_pass<0>(ms, done_cnt, tmem_d, (N + 31) & ~31, [&](int c0, uint32_t, const uint32_t (&r)[CH]) {
#pragma unroll
        for (int j4 = 0; j4 < CH / 4; j4++) {
            const int c = c0 + 4 * j4;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < N) {   // N % 4 == 0
                const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c));
                o.x = act_fast<ACT>(act, __uint_as_float(r[4 * j4]) + b.x);
                o.y = act_fast<ACT>(act, __uint_as_float(r[4 * j4 + 1]) + b.y);
                o.z = act_fast<ACT>(act, __uint_as_float(r[4 * j4 + 2]) + b.z);
                o.w = act_fast<ACT>(act, __uint_as_float(r[4 * j4 + 3]) + b.w);
            }
            *reinterpret_cast<float4 *>(XP + out_chunk_offset(row, c)) = o;
        }
    });
}
// (the last layer never has a skip connection, cpp:269-279, and GCN leaves it unscaled)
__device__ __forceinline__ void epilogue_rows(Misc &ms, uint32_t &done_cnt, uint32_t tmem_d,
                                              unsigned char *XP, int N,
                                              const float *__restrict__ bias, int act)
{
    if (act == GNNB_ACT_RELU) epilogue_rows_t<1>(ms, done_cnt, tmem_d, XP, N, bias, act);
    else if (act == GNNB_ACT_IDENTITY) epilogue_rows_t<0>(ms, done_cnt, tmem_d, XP, N, bias, act);
    else epilogue_rows_t<2>(ms, done_cnt, tmem_d, XP, N, bias, act);
}

// head output: columns [0, n_true) of rows [0, n_rows) -> gout[gids[row]][col]
__device__ __forceinline__ void epilogue_global(uint32_t tmem_d, int N, const float *__restrict__ bias,
                                                int act, float *gout, const int *gids, int n_rows,
                                                int ldg, int n_true)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = 32 * (warp & 3) + lane;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int npad = (N + 31) & ~31;
    float *grow = row < n_rows ? gout + (size_t)gids[row] * ldg : nullptr;
    for (int c0 = (warp >> 2) * 32; c0 < npad; c0 += CSTRIDE) {
        float v[32];
        tc::tmem_ld32(tmem_d + lane_base + (uint32_t)c0, v);
        if (grow == nullptr) continue;
#pragma unroll
        for (int j = 0; j < 32; j++)
            if (c0 + j < n_true) grow[c0 + j] = act_apply_compact(act, v[j] + __ldg(bias + c0 + j));
    }
}

// barrier over the 256 worker threads (the producer warp never joins it)
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(NTHREADS) : "memory"); }

#if GNNB_TC_BF2
// ---------------------------------------------------------------------------------------
// PNA (lib:1750-2157) inside the fused kernel.  Tensor memory of a PNA layer (foP = F_out rounded
// to 16, <= 96): D_id [0, foP) | D_amp [foP, 2 foP) | D_att [2 foP, 3 foP) (the A_u accumulator of
// the pre-transform aliases D_att until its rows have been copied to shared memory) | B_v, later the
// final linear's accumulator, at [3 foP, ...) | A operand at [384, 512).
__device__ __forceinline__ uint32_t pna_col_att(const PnaLayer &P) { return 2u * (uint32_t)P.foP; }
__device__ __forceinline__ uint32_t pna_col_b(const PnaLayer &P) { return 3u * (uint32_t)P.foP; }

// The in-neighbors of a row, extracted ONCE per tile from the u8 multiplicity counters into
// registers: up to 8 entries (source | multiplicity << 8, 16 bits each, ascending source order =
// deterministic).  Walking the counters inside every statistics pass instead made a warp execute
// the gather body once per (word, byte) position ANY of its lanes had a neighbor at (~27 times per
// pass, ncu: 3.2 k instructions per step); with the lists it runs max-degree-in-warp times (~4).
// n = 9 marks a row with more than 8 distinct in-neighbors: those walk the counters as before.
struct PnaNbr {
    unsigned long long lo, hi;
    int n;
};
__device__ __forceinline__ void pna_build_nbr(const uint32_t *CNT, int row, int g_r0, int g_r1, PnaNbr &nb)
{
    nb.lo = 0ull; nb.hi = 0ull; nb.n = 0;
    const uint32_t *cw = CNT + row * (TM / 4);
    for (int w = g_r0 >> 2; w < (g_r1 + 3) >> 2; w++) {
        const uint32_t word = cw[(w + row) & 31];      // (skewed layout, see the edge pass)
        if (word == 0u) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t m = (word >> (8 * b)) & 0xffu;
            if (m) {
                const unsigned long long e = (unsigned long long)((uint32_t)(4 * w + b) | (m << 8));
                if (nb.n < 4) nb.lo |= e << (16 * nb.n);
                else if (nb.n < 8) nb.hi |= e << (16 * (nb.n - 4));
                nb.n = nb.n < 9 ? nb.n + 1 : 9;
            }
        }
    }
}
template <class F>
__device__ __forceinline__ void pna_for_each_neighbor(const PnaNbr &nb, const uint32_t *CNT, int row,
                                                      int g_r0, int g_r1, F &&f)
{
    if (nb.n <= 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i < nb.n) {
                const uint32_t e = (uint32_t)((i < 4 ? nb.lo : nb.hi) >> (16 * (i & 3))) & 0xffffu;
                f((int)(e & 0xffu), (int)(e >> 8));
            }
        }
        return;
    }
    const uint32_t *cw = CNT + row * (TM / 4);
    for (int w = g_r0 >> 2; w < (g_r1 + 3) >> 2; w++) {
        uint32_t word = cw[(w + row) & 31];
        if (word == 0u) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = (int)((word >> (8 * b)) & 0xffu);
            if (m) f(4 * w + b, m);
        }
    }
}

// One PNA conv layer, worker side.  Thread (row, hh): row = TMEM lane, hh = warp >> 2 picks the
// column block inside a 64-column half like every other row pass.
__device__ __forceinline__ int pna_layer_workers(const TcParams &p, Misc &ms, int l, uint32_t tmem_base,
                                                 unsigned char *XP, float *AROWS, const uint32_t *CNT,
                                                 uint32_t &done_cnt, int my_deg, const PnaNbr &nb, int g_r0,
                                                 int g_r1, bool last_layer, bool do_skip, bool next_skip,
                                                 long long &t_prev)
{
    // GNNB_FUSED_TIMING: thread 0's cycles per phase -> timing[8 ..]
#define PNA_PHASE(idx)                                                              \
    if (p.timing != nullptr && threadIdx.x == 0) {                                  \
        const long long t_now = clock64();                                          \
        atomicAdd(p.timing + (idx), (unsigned long long)(t_now - t_prev));          \
        t_prev = t_now;                                                             \
    }
    const PnaLayer &P = p.pna[l];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = 32 * (warp & 3) + lane, hh = warp >> 2;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int fi = p.fi[l], kp = (fi + 31) & ~31;
    // ---- phase A: X rows -> A operand; the issuer runs the three GEMMs off it.  (From layer 1 on
    // the previous layer's output pass has written and handed off the A operand already.)
    if (l == 0) cvt_self(ms, tmem_base, XP, kp);
    PNA_PHASE(8)
    wait_done_both(ms, done_cnt);
    PNA_PHASE(9)
    // ---- phase B: A_u rows (fp32) -> shared memory, where every row's thread can gather them
    for (int c = 16 * hh; c < P.fiP; c += 32) {
        uint32_t r[16];
        tc::tmem_ld_nowait(tmem_base + pna_col_att(P) + lane_base + (uint32_t)c, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++)
            *reinterpret_cast<uint4 *>(AROWS + row * PNA_LDA + c + 4 * j4) =
                make_uint4(r[4 * j4], r[4 * j4 + 1], r[4 * j4 + 2], r[4 * j4 + 3]);
    }
    tc::tc_fence_before();
    worker_sync();
    tc::tc_fence_after();
    PNA_PHASE(10)
    // ---- phase C: per group of 32 features, per sub-block h of 16 features: the two threads of a
    // row take 8 features each and compute all four statistics over the row's in-neighbors (one
    // gather for max / min / sum, a second one for the variance around the mean); the 32 results
    // are A-operand columns [64 h + 32 hh, +32) of the group's GEMMs.  The weight images are
    // permuted to that order: K' = 64 h + 32 hh + 8 q + j  <->  statistic q (max, min, mean, std)
    // of feature 32 g + 16 h + 8 hh + j.
    const float degf = (float)my_deg;
    const float inv_deg = 1.0f / degf;        // (in-degree 0: inf, so the variance becomes 0 * inf = NaN, lib:702)
    for (int g = 0; g < P.ng; g++) {
        const int gw = P.gw[g];
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int f0 = 32 * g + 16 * h + 8 * hh;
            if (16 * h < gw) {
                float bv[8], vmax[8], vmin[8], vsum[8], vm2[8];
                {   // B_v + b_pre of this row (tensor memory) for its 8 features
                    uint32_t r[8];
                    tc::tmem_ld8_nowait(tmem_base + pna_col_b(P) + lane_base + (uint32_t)f0, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; j++) bv[j] = __uint_as_float(r[j]) + __ldg(P.pb.bias + f0 + j);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) { vmax[j] = 0.0f; vmin[j] = 0.0f; vsum[j] = 0.0f; vm2[j] = 0.0f; }
                // max / min: the reference's first-sample rule (lib:748-795) differs from fmaxf / fminf
                // only when a sample is NaN -- and then the mean, hence the whole output row, is NaN
                // either way -- so the running extrema start from -inf / +inf (0 for no samples)
#pragma unroll
                for (int j = 0; j < 8; j++) { vmax[j] = -INFINITY; vmin[j] = INFINITY; }
                pna_for_each_neighbor(nb, CNT, row, g_r0, g_r1, [&](int u, int m) {
                    const float *au = AROWS + u * PNA_LDA + f0;
                    const float4 a0 = *reinterpret_cast<const float4 *>(au);
                    const float4 a1 = *reinterpret_cast<const float4 *>(au + 4);
                    const float t[8] = {a0.x + bv[0], a0.y + bv[1], a0.z + bv[2], a0.w + bv[3],
                                        a1.x + bv[4], a1.y + bv[5], a1.z + bv[6], a1.w + bv[7]};
                    const float mf = (float)m;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        vmax[j] = fmaxf(vmax[j], t[j]);
                        vmin[j] = fminf(vmin[j], t[j]);
                        vsum[j] = fmaf(mf, t[j], vsum[j]);
                    }
                });
                if (my_deg == 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++) { vmax[j] = 0.0f; vmin[j] = 0.0f; }
                }
#pragma unroll
                for (int j = 0; j < 8; j++) vsum[j] = my_deg > 0 ? vsum[j] * inv_deg : 0.0f;   // mean (lib:661)
                pna_for_each_neighbor(nb, CNT, row, g_r0, g_r1, [&](int u, int m) {
                    const float *au = AROWS + u * PNA_LDA + f0;
                    const float4 a0 = *reinterpret_cast<const float4 *>(au);
                    const float4 a1 = *reinterpret_cast<const float4 *>(au + 4);
                    const float t[8] = {a0.x + bv[0], a0.y + bv[1], a0.z + bv[2], a0.w + bv[3],
                                        a1.x + bv[4], a1.y + bv[5], a1.z + bv[6], a1.w + bv[7]};
                    const float mf = (float)m;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float d = t[j] - vsum[j];
                        vm2[j] = fmaf(mf * d, d, vm2[j]);
                    }
                });
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float var = vm2[j] * inv_deg + 1e-5f;      // population variance (lib:702-703)
                    v[j] = vmax[j]; v[8 + j] = vmin[j]; v[16 + j] = vsum[j];
                    v[24 + j] = var * rsqrtf(var);                    // sqrt(var), 2 ulp; NaN for in-degree 0
                }
                PNA_PHASE(11)
                // the MMAs that read THIS half of the A operand must be done: the previous group's
                // (the issuer commits done[h] per half: atom-major order), for group 0 the S GEMM
                tc::mbar_wait(&ms.bar_done[h], done_cnt & 1);
                tc::tc_fence_after();
                PNA_PHASE(12)
                split_store(tmem_base, lane_base, 64 * h + 32 * hh, v);
                tc::tmem_st_wait();
            } else {
                tc::mbar_wait(&ms.bar_done[h], done_cnt & 1);
            }
            handoff_half(ms, h);
            PNA_PHASE(13)
        }
        done_cnt++;
    }
    wait_done_both(ms, done_cnt);
    PNA_PHASE(12)
    // ---- phase D: S + D_id + amp_v D_amp + att_v D_att + b_post -> A operand of the final linear
    {
        const float lg = logf((float)(my_deg < 1 ? 1 : my_deg) + 1.0f);      // lib:1973-1984
        const float amp = lg / p.pna_delta, att = p.pna_delta / lg;
        const int fo = p.fo[l], npad = (fo + 31) & ~31;
#pragma unroll 1
        for (int st = 0; st < PASS_STEPS; st++) {
            const int c0 = 64 * st + 32 * hh;
            if (c0 < npad) {
                float v[32];
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int c = c0 + 16 * q;
                    if (c < fo) {   // fo % 16 == 0
                        uint32_t a[16], b[16], d[16];
                        tc::tmem_ld_nowait(tmem_base + lane_base + (uint32_t)c, a);
                        tc::tmem_ld_nowait(tmem_base + (uint32_t)P.foP + lane_base + (uint32_t)c, b);
                        tc::tmem_ld_nowait(tmem_base + pna_col_att(P) + lane_base + (uint32_t)c, d);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            v[16 * q + j] = fmaf(att, __uint_as_float(d[j]),
                                                 fmaf(amp, __uint_as_float(b[j]),
                                                      __uint_as_float(a[j]) + __ldg(P.b_post + c + j)));
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j++) v[16 * q + j] = 0.0f;
                    }
                }
                split_store(tmem_base, lane_base, c0, v);
            }
            tc::tmem_st_wait();
            handoff_half(ms, st);
        }
    }
    PNA_PHASE(14)
    // ---- phase E: final linear's accumulator -> (+bias, skip, activation) -> next layer / pooling
    const uint32_t tmem_d = tmem_base + pna_col_b(P);
    if (last_layer) {
        epilogue_rows(ms, done_cnt, tmem_d, XP, P.pl.N, P.pl.bias, p.gnn_act);
        return 0;
    }
    // the planes are only read back as the next layer's skip source (by this very thread): no
    // shared-memory fence, no barrier; the A operand is handed off per 64-column half
    (void)next_skip;
    if (p.gnn_act == GNNB_ACT_RELU) {
        if (do_skip) epilogue_planes_a_t<1, true>(ms, done_cnt, tmem_base, tmem_d, XP, P.pl.N, P.pl.bias, p.gnn_act);
        else epilogue_planes_a_t<1, false>(ms, done_cnt, tmem_base, tmem_d, XP, P.pl.N, P.pl.bias, p.gnn_act);
    } else {
        if (do_skip) epilogue_planes_a_t<2, true>(ms, done_cnt, tmem_base, tmem_d, XP, P.pl.N, P.pl.bias, p.gnn_act);
        else epilogue_planes_a_t<2, false>(ms, done_cnt, tmem_base, tmem_d, XP, P.pl.N, P.pl.bias, p.gnn_act);
    }
    return 0;
#undef PNA_PHASE
}

// The three GEMMs of one feature group (D_id += s.W_id^T, D_amp (+)= s.W_amp^T, D_att (+)= s.W_att^T)
// issued ATOM-MAJOR: all MMAs that read A-operand half 0 (K' [0, 64)) first, then done[0] is
// committed, then half 1 -- so the workers may overwrite half 0 with the next group's statistics
// while the tensor pipe is still busy with half 1.  Unit order in the weight stream: per atom,
// per linear (id, amp, att): hi unit, mid unit.
__device__ __forceinline__ void pna_group_issue(Misc &ms, uint32_t ring, uint32_t &cons, uint32_t tmem_base,
                                                const PnaLayer &P, int g, uint32_t &ready_cnt)
{
    const bool leader = tc::elect_one();
    const uint32_t slot_log2 = ms.slot_log2, slot_stride = ms.slot_stride;
    const uint32_t idesc = tc::make_idesc_bf16(TM, P.foP, 0);
    const uint32_t ahi = tmem_base + TM_AHI, amid = tmem_base + TM_ALO;
    const int KA = P.gid[g].KA;          // 1 (16-feature group, K' = 64) or 2
    for (int ka = 0; ka < 2; ka++) {
        tc::mbar_wait(&ms.bar_ready[ka], ready_cnt & 1);
        tc::tc_fence_after();
        if (ka < KA) {
            const uint32_t col = (uint32_t)(ka * 32);
#pragma unroll 1
            for (int lin = 0; lin < 3; lin++) {
                const uint32_t tmem_d = tmem_base + (uint32_t)(lin * P.foP);
                // D_id holds the self block already; D_amp / D_att start with the first group
                const uint32_t first = (lin == 0 || g > 0 || ka > 0) ? 1u : 0u;
                for (int part = 0; part < 2; part++) {     // hi unit, then mid unit
                    const uint32_t sl = cons & ((1u << slot_log2) - 1u);
                    tc::mbar_wait(&ms.bar_full[sl], (cons >> slot_log2) & 1);
                    tc::tc_fence_after();
                    const uint64_t bd = tc::make_desc(ring + sl * slot_stride);
                    if (leader) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; k4++) {
                            if (part == 0) {
                                tc::mma_bf16_ts(tmem_d, ahi + col + 8 * k4, bd + (uint64_t)(2 * k4), idesc,
                                                k4 == 0 ? first : 1u);
                                tc::mma_bf16_ts(tmem_d, amid + col + 8 * k4, bd + (uint64_t)(2 * k4), idesc, 1u);
                            } else {
                                tc::mma_bf16_ts(tmem_d, ahi + col + 8 * k4, bd + (uint64_t)(2 * k4), idesc, 1u);
                            }
                        }
                        tc::mma_commit(&ms.bar_empty[sl]);
                    }
                    cons++;
                }
            }
        }
        if (leader) tc::mma_commit(&ms.bar_done[ka]);
    }
    ready_cnt++;
}
// the weight units of one group in that order (producer warp)
__device__ __forceinline__ void pna_group_produce(Misc &ms, uint32_t ring, uint32_t &prod, const PnaLayer &P,
                                                  int g, size_t copy_off, bool leader)
{
    const uint32_t slot_log2 = ms.slot_log2, slot_stride = ms.slot_stride;
    const uint32_t bytes = (uint32_t)P.foP * tc::ROW_BYTES;
    const int KA = P.gid[g].KA;
    for (int ka = 0; ka < KA; ka++)
        for (int lin = 0; lin < 3; lin++) {
            const TLinear &L = lin == 0 ? P.gid[g] : lin == 1 ? P.gamp[g] : P.gatt[g];
            const unsigned char *src = reinterpret_cast<const unsigned char *>(L.img) + copy_off +
                                       (size_t)ka * 2 * bytes;
            for (int part = 0; part < 2; part++) {
                const uint32_t sl = prod & ((1u << slot_log2) - 1u), use = prod >> slot_log2;
                if (use > 0) tc::mbar_wait(&ms.bar_empty[sl], (use - 1) & 1);
                if (leader) {
                    tc::mbar_expect_tx(&ms.bar_full[sl], bytes);
                    tc::bulk_g2s_addr(ring + sl * slot_stride, src + (size_t)part * bytes, bytes, &ms.bar_full[sl]);
                }
                prod++;
            }
        }
}

// the same layer on the MMA-issuing warp
__device__ __forceinline__ void pna_layer_issue(const TcParams &p, Misc &ms, int l, uint32_t ring,
                                                uint32_t &cons, uint32_t tmem_base, uint32_t &ready_cnt)
{
    const PnaLayer &P = p.pna[l];
    const uint32_t c_att = 2u * (uint32_t)P.foP, c_b = 3u * (uint32_t)P.foP;
    // A_u and B_v first (the workers wait for these two), S as a phase of its own: it is only needed
    // by the combine pass, so it runs while the workers copy A_u and gather group 0's statistics
    gemm_issue(ms, ring, cons, tmem_base, c_att, P.pa, false, ready_cnt, true, false);
    gemm_issue(ms, ring, cons, tmem_base, c_b, P.pb, false, ready_cnt, false, true);
    gemm_issue(ms, ring, cons, tmem_base, 0u, P.ps, false, ready_cnt, false, true);
    for (int g = 0; g < P.ng; g++) pna_group_issue(ms, ring, cons, tmem_base, P, g, ready_cnt);
    gemm_issue(ms, ring, cons, tmem_base, c_b, P.pl, false, ready_cnt);
}
// ... and the weight units it consumes, in the same order (producer warp)
__device__ __forceinline__ void pna_layer_produce(const TcParams &p, Misc &ms, int l, uint32_t ring,
                                                  uint32_t &prod, size_t copy_off, bool leader)
{
    const PnaLayer &P = p.pna[l];
    produce_linear(ms, ring, prod, P.pa, copy_off, leader);
    produce_linear(ms, ring, prod, P.pb, copy_off, leader);
    produce_linear(ms, ring, prod, P.ps, copy_off, leader);
    for (int g = 0; g < P.ng; g++) pna_group_produce(ms, ring, prod, P, g, copy_off, leader);
    produce_linear(ms, ring, prod, P.pl, copy_off, leader);
}
#endif   // GNNB_TC_BF2

// MLP head (cpp:454-530) for up to 128 pending graphs: the pooled vectors [128][head_in] go from the
// per-CTA pending buffer (L2) to tensor memory in 128-wide K chunks that accumulate into the same
// accumulator; later head layers take their A operand from the previous epilogue like GIN's hidden
// layer.  All worker threads call this.  dw: index of the accumulator buffer of the latest
// non-accumulating MMA phase (toggled whenever the workers start one).
__device__ __forceinline__ void head_flush(const TcParams &p, Misc &ms, uint32_t tmem_base,
                                           const float *pending, int n_rows, uint32_t &done_cnt,
                                           uint32_t &dw)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = 32 * (warp & 3) + lane;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int head_in = p.emb * p.num_pools;
    for (int j = 0; j < p.mlp_num_linear; j++) {
        const bool last = j == p.mlp_num_linear - 1;
        const int nch = p.hchunks[j];
        for (int c = 0; c < nch; c++) {
            const TLinear &L = p.hl[j][c];
            if (j == 0) {  // A chunk: pending[:, 128c : 128c + K) -> (hi, lo), zero padded
                const int kp = (L.K + 31) & ~31;
                const float *src = pending + (size_t)row * PLD + c * 128;
#pragma unroll 1
                for (int st = 0; st < PASS_STEPS; st++) {
                    const int h = st, c0 = 64 * st + STEP_COLS * (warp >> 2);
                    if (c0 < kp) {
#pragma unroll 1
                        for (int cc = c0; cc < c0 + STEP_COLS; cc += CH) {
                            float v[CH];
#pragma unroll
                            for (int j4 = 0; j4 < CH / 4; j4++) {
                                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (row < n_rows && c * 128 + cc + j4 * 4 < head_in)   // head_in % 4 == 0
                                    t = __ldcg(reinterpret_cast<const float4 *>(src + cc + j4 * 4));
                                v[j4 * 4] = t.x; v[j4 * 4 + 1] = t.y; v[j4 * 4 + 2] = t.z; v[j4 * 4 + 3] = t.w;
                            }
                            split_store(tmem_base, lane_base, cc, v);
                        }
                    }
                    tc::tmem_st_wait();
                    handoff_half(ms, h);     // (for j > 0 the previous layer's epilogue handed A off)
                }
                if (c == 0) dw ^= 1u;        // chunk 0 starts a new accumulator, the others add to it
            }
            // the next chunk overwrites the A operand: wait for the MMAs that read it
            if (c < nch - 1) wait_done_both(ms, done_cnt);
        }
        const TLinear &L0 = p.hl[j][0];
        if (last) {
            wait_done_both(ms, done_cnt);
            epilogue_global(tmem_base + dcol_of(dw), L0.N, L0.bias, p.out_act, p.out, ms.pend_gid, n_rows,
                            p.mlp_out, p.head_n[j]);
        } else {
            epilogue_tmem(ms, done_cnt, tmem_base, dcol_of(dw), L0.N, L0.bias, p.mlp_act);
            dw ^= 1u;                        // ... which started the next head layer's GEMM
        }
    }
}

// the head's GEMMs in execution order (issuing warp)
__device__ __forceinline__ void head_issue(const TcParams &p, Misc &ms, uint32_t ring, uint32_t &cons,
                                           uint32_t tmem_base, uint32_t &ready_cnt, uint32_t &dw)
{
    for (int j = 0; j < p.mlp_num_linear; j++)
        for (int c = 0; c < p.hchunks[j]; c++) {
            if (c == 0) dw ^= 1u;
            gemm_issue(ms, ring, cons, tmem_base, dcol_of(dw), p.hl[j][c], c > 0, ready_cnt);
        }
}

__global__ void __launch_bounds__(CTA_THREADS, 1) fused_tc_kernel(const __grid_constant__ TcParams p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *ADJ = base;                   // (PNA: the A_u rows, fp32 [128][PNA_LDA])
    unsigned char *XP = ADJ + p.r0_bytes;
    unsigned char *RING = XP + NPLANES * tc::PLANE_BYTES;
    uint32_t *CNT = reinterpret_cast<uint32_t *>(RING + p.ring_bytes);
    Misc &ms = *reinterpret_cast<Misc *>(RING + p.ring_bytes + CNT_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_prev = clock64();
#define GNNB_PHASE(idx)                                                             \
    if (p.timing != nullptr && tid == 0) {                                          \
        const long long t_now = clock64();                                          \
        atomicAdd(p.timing + (idx), (unsigned long long)(t_now - t_prev));          \
        t_prev = t_now;                                                             \
    }

    if (warp == 0) tc::tmem_alloc(&ms.tmem_slot, TMEM_COLS);
    if (tid == 0) {
        ms.slot_log2 = (uint32_t)p.slot_log2;
        ms.slot_stride = (uint32_t)p.slot_stride;
        ms.timing = p.timing;
        for (int i = 0; i < MAX_NSLOT; i++) {
            tc::mbar_init(&ms.bar_full[i], 1);
            tc::mbar_init(&ms.bar_empty[i], 1);
        }
        for (int h = 0; h < 2; h++) {
            tc::mbar_init(&ms.bar_done[h], 1);
            tc::mbar_init(&ms.bar_ready[h], READY_ARRIVALS);   // one arrival per worker warp and half
        }
        tc::mbar_fence_init();
        ms.nonfinite = 0;
    }
    for (int i = tid; i < CNT_BYTES / 16; i += CTA_THREADS)
        reinterpret_cast<uint4 *>(CNT)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    // warp-uniform copies (broadcast from lane 0) of everything the issuing warp touches
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, ms.tmem_slot, 0);
    const uint32_t adj_addr = __shfl_sync(0xffffffffu, tc::smem_u32(ADJ), 0);
    const uint32_t xp_addr = adj_addr + (uint32_t)p.r0_bytes;
    const uint32_t ring_addr = xp_addr + NPLANES * tc::PLANE_BYTES;
    uint32_t done_cnt = 0, cons = 0, dw = 0;   // dw: accumulator buffer of the latest new MMA phase
    const int n_tiles = __shfl_sync(0xffffffffu, __ldg(p.n_tiles_ptr), 0);

    if (warp_u == NWARPS) {
        // ------------------------------------------------------------ weight-producer warp
        // mirrors the workers' control flow (which tiles run, when the head runs) from the tile
        // geometry alone and streams the weight units of every linear in execution order.
        // All CTAs stream the same weights at about the same time: each uses one of a few
        // replicas at different addresses so that the reads spread over the L2 slices.
        const bool leader = tc::elect_one();
        const size_t copy_off = (size_t)(blockIdx.x % p.img_copies) * p.img_copy_bytes;
        uint32_t prod = 0;
        int pend = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int tg0 = __ldg(p.tile_bounds + tile), tg1 = __ldg(p.tile_bounds + tile + 1);
            const int tng = tg1 - tg0;
            const int64_t trows = __ldg(p.node_ptr + tg1) - __ldg(p.node_ptr + tg0);
            if (tng <= 0 || trows > TM || tng > TM) continue;
            for (int l = 0; l < p.num_layers; l++) {
#if GNNB_TC_BF2
                if (p.conv_type == GNNB_CONV_PNA) {
                    pna_layer_produce(p, ms, l, ring_addr, prod, copy_off, leader);
                    continue;
                }
#endif
                produce_linear(ms, ring_addr, prod, p.l0[l], copy_off, leader);
                if (p.conv_type != GNNB_CONV_GCN) produce_linear(ms, ring_addr, prod, p.l1[l], copy_off, leader);
            }
            pend += tng;
            if (pend >= HEAD_G) {
                for (int j = 0; j < p.mlp_num_linear; j++)
                    for (int c = 0; c < p.hchunks[j]; c++)
                        produce_linear(ms, ring_addr, prod, p.hl[j][c], copy_off, leader);
                pend -= HEAD_G;
            }
        }
        if (pend > 0)
            for (int j = 0; j < p.mlp_num_linear; j++)
                for (int c = 0; c < p.hchunks[j]; c++)
                    produce_linear(ms, ring_addr, prod, p.hl[j][c], copy_off, leader);
    } else if (warp_u == NWARPS + 1) {
        // ------------------------------------------------------------ MMA-issuing warp
        // mirrors the workers' sequence of MMA phases (per tile: aggregation and transforms of
        // every layer, then the head when 128 pooled graphs are waiting); each phase consumes one
        // completion of ready[0] and ready[1] and commits once to done[0] and done[1].
        uint32_t ready_cnt = 0, dw = 0;
        int pend = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int tg0 = __ldg(p.tile_bounds + tile), tg1 = __ldg(p.tile_bounds + tile + 1);
            const int tng = tg1 - tg0;
            const int64_t trows = __ldg(p.node_ptr + tg1) - __ldg(p.node_ptr + tg0);
            if (tng <= 0 || trows > TM || tng > TM) continue;
            for (int l = 0; l < p.num_layers; l++) {
#if GNNB_TC_BF2
                if (p.conv_type == GNNB_CONV_PNA) {
                    pna_layer_issue(p, ms, l, ring_addr, cons, tmem_base, ready_cnt);
                    continue;
                }
#endif
                dw ^= 1u;
                agg_issue(ms, tmem_base + dcol_of(dw), adj_addr, xp_addr, (p.fi[l] + 31) & ~31, (int)trows,
                          ready_cnt);
                dw ^= 1u;
                gemm_issue(ms, ring_addr, cons, tmem_base, dcol_of(dw), p.l0[l], false, ready_cnt);
                if (p.conv_type != GNNB_CONV_GCN) {
                    const bool acc = p.conv_type == GNNB_CONV_SAGE;
                    if (!acc) dw ^= 1u;
                    gemm_issue(ms, ring_addr, cons, tmem_base, dcol_of(dw), p.l1[l], acc, ready_cnt);
                }
            }
            pend += tng;
            if (pend >= HEAD_G) {
                head_issue(p, ms, ring_addr, cons, tmem_base, ready_cnt, dw);
                pend -= HEAD_G;
            }
        }
        if (pend > 0) head_issue(p, ms, ring_addr, cons, tmem_base, ready_cnt, dw);
    } else {
    // ---------------------------------------------------------------- worker warps
    float *pending = p.pending + (size_t)blockIdx.x * HEAD_G * PLD;
    const int emb = p.emb;
    const int conv = p.conv_type;
    const bool self_loop = conv == GNNB_CONV_GCN || conv == GNNB_CONV_GIN;
    int bad_values = 0;
    int pend_n = 0;   // pooled graphs waiting for the head (uniform over the workers)

    // geometry of the first tile; later tiles are prefetched one iteration ahead
    int g0 = 0, g1 = 0;
    int64_t row0 = 0, e0 = 0, row1 = 0, e1 = 0;
    if ((int)blockIdx.x < n_tiles) {
        g0 = __ldg(p.tile_bounds + blockIdx.x);
        g1 = __ldg(p.tile_bounds + blockIdx.x + 1);
        row0 = __ldg(p.node_ptr + g0); row1 = __ldg(p.node_ptr + g1);
        e0 = __ldg(p.edge_ptr + g0); e1 = __ldg(p.edge_ptr + g1);
    }

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int ng = g1 - g0;
        const int rows = (int)(row1 - row0);
        const int ne = (int)(e1 - e0);
        const int cur_g0 = g0;
        const int64_t cur_row0 = row0, cur_e0 = e0;
        const int nxt = tile + gridDim.x;
        int ng0 = 0, ng1 = 0;
        if (nxt < n_tiles) {
            ng0 = __ldg(p.tile_bounds + nxt);
            ng1 = __ldg(p.tile_bounds + nxt + 1);
        }
        const bool bad = rows > TM || ng > TM;
        if (bad && tid == 0) atomicExch(p.error_flag, 1);
        const bool run = ng > 0 && !bad;
        const int r_own = tid & (TM - 1), c_half = tid >> 7;   // staging role: (row, chunk parity)
        const int kp0 = (p.in_dim + 31) & ~31;   // plane columns staged (aggregation K atoms of 32 features)
        if (run) {
            worker_sync();   // the previous tile's pooling has finished reading the planes
            // ---------------------------------------------------------------- stage inputs
            for (int i = tid; i <= ng; i += NTHREADS) {
                ms.grow[i] = (int)(__ldg(p.node_ptr + cur_g0 + i) - cur_row0);
                ms.gedge[i] = (int)(__ldg(p.edge_ptr + cur_g0 + i) - cur_e0);
            }
            if (tid < TM) ms.deg[tid] = 0;
            // first two feature chunks of this thread's row: loads in flight during the edge pass
            float xv[2][8];
            {
                const int F = p.in_dim;
                const float *src = p.x + (size_t)(cur_row0 + r_own) * F;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int c = (c_half + NCHALF * q) * 8;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        xv[q][j] = (r_own < rows && c + j < F) ? __ldg(src + c + j) : 0.0f;
                }
            }
            // first edge of this thread: in flight across the barrier as well
            const int2 *coo = reinterpret_cast<const int2 *>(p.coo) + cur_e0;
            const int2 first_edge = tid < ne ? __ldg(coo + tid) : make_int2(0, 0);
            worker_sync();
            {   // edges -> multiplicity counts + in-degrees (lib:1051-1083)
                for (int j = tid; j < ne; j += NTHREADS) {
                    const int2 sd = j == tid ? first_edge : __ldg(coo + j);
                    int lo = 0, hi = ng;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (ms.gedge[mid] <= j) lo = mid; else hi = mid;
                    }
                    const int b = ms.grow[lo], n_i = ms.grow[lo + 1] - b;
                    if ((unsigned)sd.x >= (unsigned)n_i || (unsigned)sd.y >= (unsigned)n_i) {
                        atomicExch(p.error_flag, 2);
                    } else {
                        const int ls = sd.x + b, ld = sd.y + b;
                        const uint32_t sh = 8u * (uint32_t)(ls & 3);
                        // PNA reads the counters row by row (thread per row): word w of row d sits at
                        // position (w + d) & 31 so that the 32 rows of a warp hit 32 different banks
                        const int wpos = conv == GNNB_CONV_PNA ? (((ls >> 2) + ld) & 31) : (ls >> 2);
                        const uint32_t old = atomicAdd(&CNT[ld * (TM / 4) + wpos], 1u << sh);
                        if (((old >> sh) & 0xffu) == 0xffu) atomicExch(p.error_flag, 1);
                        atomicAdd(&ms.deg[ld], 1);
                    }
                }
            }
            worker_sync();
            GNNB_PHASE(0)
            // ---------------------------------------------------------------- ADJ, planes
            // (PNA gathers over the multiplicity counters themselves: no ADJ, counters kept)
            for (int idx = tid; idx < (conv == GNNB_CONV_PNA ? 0 : TM * 8); idx += NTHREADS) {   // 16 sources per iteration
                const int d = idx >> 3, s0 = (idx & 7) * 16;
                uint4 *cp = reinterpret_cast<uint4 *>(CNT) + idx;
                const uint4 cw = *cp;
                // most 16-source groups of the block-diagonal matrix are empty: zeros, no conversion
                if ((cw.x | cw.y | cw.z | cw.w) == 0u &&
                    !(self_loop && d < rows && (unsigned)(d - s0) < 16u)) {
                    *reinterpret_cast<uint4 *>(ADJ + tc::adj_chunk_offset(d, s0)) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4 *>(ADJ + tc::adj_chunk_offset(d, s0 + 8)) = make_uint4(0u, 0u, 0u, 0u);
                    continue;
                }
                *cp = make_uint4(0u, 0u, 0u, 0u);
                const uint32_t w4[4] = {cw.x, cw.y, cw.z, cw.w};
                uint32_t o[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float f[4];
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int s = s0 + 4 * q + b;
                        f[b] = (float)((w4[q] >> (8 * b)) & 0xffu) +
                               ((self_loop && s == d && d < rows) ? 1.0f : 0.0f);
                    }
                    o[2 * q] = __byte_perm(__float_as_uint(f[0]), __float_as_uint(f[1]), 0x7632);
                    o[2 * q + 1] = __byte_perm(__float_as_uint(f[2]), __float_as_uint(f[3]), 0x7632);
                }
                *reinterpret_cast<uint4 *>(ADJ + tc::adj_chunk_offset(d, s0)) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4 *>(ADJ + tc::adj_chunk_offset(d, s0 + 8)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            {   // node features -> planes (GCN: pre-scaled by dinv), zero padded to whole K atoms
                const int F = p.in_dim;
                const float sc = conv == GNNB_CONV_GCN ? 1.0f / sqrtf(1.0f + (float)ms.deg[r_own]) : 1.0f;
                const float *src = p.x + (size_t)(cur_row0 + r_own) * F;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int c = (c_half + NCHALF * q) * 8;
                    if (c < kp0) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            o[j] = xv[q][j] * sc;
                            bad_values |= ((__float_as_uint(o[j]) & 0x7f800000u) == 0x7f800000u) ? 1 : 0;
                        }
                        store_row8(XP, r_own, c, o);
                    }
                }
                for (int c = (c_half + 2 * NCHALF) * 8; c < kp0; c += 8 * NCHALF) {   // in_dim > 32 (16 NCHALF)
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        o[j] = ((r_own < rows && c + j < F) ? __ldg(src + c + j) : 0.0f) * sc;
                        bad_values |= ((__float_as_uint(o[j]) & 0x7f800000u) == 0x7f800000u) ? 1 : 0;
                    }
                    store_row8(XP, r_own, c, o);
                }
            }
            if (conv == GNNB_CONV_PNA) {
                bad_values = 0;      // PNA never mixes the graphs of a tile: NaN / Inf stay in their graph
                worker_sync();       // layer 0's row pass reads plane chunks other threads staged
            } else {
                tc::fence_async_smem();
                handoff_both(ms);    // ADJ + planes are ready: layer 0's aggregation may start
                dw ^= 1u;
            }
            GNNB_PHASE(1)
        }
        // second half of the next tile's geometry (the bounds have arrived by now), and a hint to
        // pull that tile's features / edges / offsets into L2 while this tile computes
        if (nxt < n_tiles) {
            g0 = ng0; g1 = ng1;
            row0 = __ldg(p.node_ptr + ng0); row1 = __ldg(p.node_ptr + ng1);
            e0 = __ldg(p.edge_ptr + ng0); e1 = __ldg(p.edge_ptr + ng1);
            const char *xb = reinterpret_cast<const char *>(p.x + (size_t)row0 * p.in_dim);
            const char *eb = reinterpret_cast<const char *>(p.coo + 2 * (size_t)e0);
            const size_t xbytes = (size_t)(row1 - row0) * p.in_dim * 4, ebytes = (size_t)(e1 - e0) * 8;
            for (size_t o = (size_t)tid * 128; o < xbytes; o += (size_t)NTHREADS * 128) tc::prefetch_l2(xb + o);
            for (size_t o = (size_t)tid * 128; o < ebytes; o += (size_t)NTHREADS * 128) tc::prefetch_l2(eb + o);
        }
        if (!run) continue;

        const int my_deg = ms.deg[32 * (warp & 3) + lane];   // row owned in the thread-per-row phases
        const float my_dinv = 1.0f / sqrtf(1.0f + (float)my_deg);

        // -------------------------------------------------------------------- conv layers
#if GNNB_TC_BF2
        if (conv == GNNB_CONV_PNA) {
            // the row range of this thread's graph (rows beyond the tile's graphs: empty range)
            const int my_row = 32 * (warp & 3) + lane;
            int g_r0 = 0, g_r1 = 0;
            if (my_row < rows) {
                int lo = 0, hi = ng;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (ms.grow[mid] <= my_row) lo = mid; else hi = mid;
                }
                g_r0 = ms.grow[lo];
                g_r1 = ms.grow[lo + 1];
            }
            PnaNbr nb;
            pna_build_nbr(CNT, my_row, g_r0, g_r1, nb);
            for (int l = 0; l < p.num_layers; l++) {
                const bool last_layer = l == p.num_layers - 1;
                pna_layer_workers(p, ms, l, tmem_base, XP, reinterpret_cast<float *>(ADJ), CNT, done_cnt,
                                  my_deg, nb, g_r0, g_r1, last_layer, p.skip && l != 0 && !last_layer,
                                  p.skip && l + 1 < p.num_layers - 1, t_prev);
                GNNB_PHASE(7)
            }
            worker_sync();           // pooling reads the other threads' rows; the counters are dead
            for (int i = tid; i < CNT_BYTES / 16; i += NTHREADS)
                reinterpret_cast<uint4 *>(CNT)[i] = make_uint4(0u, 0u, 0u, 0u);
        } else
#endif
        for (int l = 0; l < p.num_layers; l++) {
            const int fi = p.fi[l];
            // A-operand columns the row passes define; the MMAs read whole k-steps of real K only
            const int kp = (fi + 31) & ~31;
            const bool last_layer = l == p.num_layers - 1;
            const bool do_skip = p.skip && l != 0 && !last_layer;  // cpp:269-279
            // aggregation accumulator -> A operand (each half as soon as its MMAs are done), which
            // starts the first transform
            if (conv == GNNB_CONV_GCN) cvt_agg<0>(ms, done_cnt, tmem_base, dcol_of(dw), XP, kp, my_dinv, 0.0f);
            else if (conv == GNNB_CONV_SAGE)   // mean = sum * (1 / deg): one division per row, not per element
                cvt_agg<0>(ms, done_cnt, tmem_base, dcol_of(dw), XP, kp, my_deg > 0 ? 1.0f / (float)my_deg : 0.0f, 0.0f);
            else if (p.gin_eps != 0.0f) cvt_agg<2>(ms, done_cnt, tmem_base, dcol_of(dw), XP, kp, 1.0f, p.gin_eps);
            else cvt_agg<1>(ms, done_cnt, tmem_base, dcol_of(dw), XP, kp, 1.0f, 0.0f);
            dw ^= 1u;
            GNNB_PHASE(2)
            if (conv == GNNB_CONV_GIN) {
                epilogue_tmem(ms, done_cnt, tmem_base, dcol_of(dw), p.l0[l].N, p.l0[l].bias, GNNB_ACT_RELU);
                dw ^= 1u;
                GNNB_PHASE(3)
            } else if (conv == GNNB_CONV_SAGE) {
                wait_done_both(ms, done_cnt);     // the first transform has finished reading A
                cvt_self(ms, tmem_base, XP, kp);  // the root transform accumulates on the same buffer
                GNNB_PHASE(3)
            }
            {
                const TLinear &Lb = conv == GNNB_CONV_GIN ? p.l1[l] : p.l0[l];
                const bool gcn = conv == GNNB_CONV_GCN;
                if (last_layer) {   // only pooling reads it: plain fp32 rows
                    epilogue_rows(ms, done_cnt, tmem_base + dcol_of(dw), XP, Lb.N, Lb.bias, p.gnn_act);
                    worker_sync();           // pooling reads the other threads' rows
                } else {            // the planes of the next layer; each half starts its aggregation
                    bad_values |= epilogue_planes(ms, done_cnt, tmem_base + dcol_of(dw), XP, Lb.N, Lb.bias,
                                                  p.gnn_act, do_skip, gcn, sqrtf(1.0f + (float)my_deg), my_dinv);
                    dw ^= 1u;
                }
            }
            GNNB_PHASE(7)
        }

        // -------------------------------------------------------------------- pooling -> pending
        for (int gdone = 0; gdone < ng;) {
            const int base_n = pend_n;     // (a register: every worker tracks the same count)
            const int cnt = min(HEAD_G - base_n, ng - gdone);
            for (int gi = warp; gi < cnt; gi += NWARPS) {
                const int r0 = ms.grow[gdone + gi], r1 = ms.grow[gdone + gi + 1];
                float *dst = pending + (size_t)(base_n + gi) * PLD;
                for (int cc = lane * 4; cc < emb; cc += 128) {   // emb % 16 == 0
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), mx = sum;
                    // rows in order (the reference's sum order, lib:2709-2739), loads of four rows
                    // in flight at a time
                    for (int r = r0; r < r1; r += 4) {
                        float4 v[4];
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            v[q] = *reinterpret_cast<const float4 *>(XP + out_chunk_offset(min(r + q, r1 - 1), cc));
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            if (r + q < r1) {
                                sum.x += v[q].x; sum.y += v[q].y; sum.z += v[q].z; sum.w += v[q].w;
                                const bool first = r + q == r0;                       // lib:748-759
                                mx.x = (first || v[q].x > mx.x) ? v[q].x : mx.x;
                                mx.y = (first || v[q].y > mx.y) ? v[q].y : mx.y;
                                mx.z = (first || v[q].z > mx.z) ? v[q].z : mx.z;
                                mx.w = (first || v[q].w > mx.w) ? v[q].w : mx.w;
                            }
                        }
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
            pend_n = base_n + cnt;
            gdone += cnt;
            GNNB_PHASE(4)
            if (pend_n == HEAD_G) {
                worker_sync();       // the pooled vectors and graph ids of every warp are in place
                head_flush(p, ms, tmem_base, pending, HEAD_G, done_cnt, dw);
                pend_n = 0;
                GNNB_PHASE(5)
                // more graphs of this tile follow: their pooling overwrites the pending rows and ids
                // the slowest threads may still be reading in the head's last epilogue
                if (gdone < ng) worker_sync();
            }
        }
    }
    // graphs still waiting for the head
    worker_sync();
    if (pend_n > 0) head_flush(p, ms, tmem_base, pending, pend_n, done_cnt, dw);
    GNNB_PHASE(5)
#undef GNNB_PHASE
    if (bad_values) atomicExch(p.error_flag, 3);
    }   // worker warps
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

struct TcPlan {
    TcParams params{};
    DeviceBuf images, bounds, pack_tmp, flag, pending, timing;
    size_t smem_bytes = 0;
};

int fused_tc_prepare(gnnb_model *m)
{
    m->fused_tc = nullptr;
    const gnnb_model_desc &d = m->d;
    if (getenv("GNNB_DISABLE_TC") != nullptr) return GNNB_OK;
    const bool is_pna = d.conv_type == GNNB_CONV_PNA;
    const bool conv_ok = d.conv_type == GNNB_CONV_GCN || d.conv_type == GNNB_CONV_GIN ||
                         d.conv_type == GNNB_CONV_SAGE || (is_pna && BF2);
    const int head_in = m->emb_dim() * d.num_pools;
    bool dims_ok = d.num_layers >= 1 && d.num_layers <= MAX_LAYERS && d.mlp_num_linear >= 1 &&
                   d.mlp_num_linear <= MAX_HEAD && d.in_dim <= MAX_DIM && d.mlp_hidden <= MAX_DIM &&
                   d.mlp_out <= MAX_DIM && head_in <= 512;
    for (int k = 0; k < d.num_layers && dims_ok; k++) {
        int fi, fo;
        m->layer_dims(k, &fi, &fo);
        dims_ok = fo % 16 == 0 && fo >= 16 && fo <= MAX_DIM && fi <= MAX_DIM;
        if (is_pna) {   // tensor-memory budget of a PNA layer: 3 foP + max(fiP, foP) <= 384 columns
            const int fiP = (fi + 15) / 16 * 16;
            // (the A_u accumulator of the pre-transform aliases D_att: F_in may not exceed F_out)
            dims_ok = dims_ok && fo <= 96 && fiP <= fo && 4 * fo <= 384 &&
                      d.num_layers <= MAX_PNA_LAYERS && d.pna_delta > 0.0f;
        }
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
    auto build_image = [](const float *W, int N, int n_valid, int K, int ld, int col0,
                          std::vector<float> &out) {
        if (BF2) build_weight_image_bf2(W, N, n_valid, K, ld, col0, out);
        else build_weight_image(W, N, n_valid, K, ld, col0, out);
    };
    struct Pending { size_t off; int K, N; };
    std::vector<Pending> pend;
    struct PnaPending {
        struct Lin { size_t off; int K, N; size_t bias; bool has_bias; };
        Lin pa, pb, ps, gid[MAX_PNA_GROUPS], gamp[MAX_PNA_GROUPS], gatt[MAX_PNA_GROUPS], pl;
        size_t b_post;
        int ng, gw[MAX_PNA_GROUPS], fiP, foP;
    };
    std::vector<PnaPending> pna_pend;
    size_t idx = 2 * (size_t)d.mlp_num_linear;
    for (int k = 0; k < d.num_layers; k++) {
        int fi, fo;
        m->layer_dims(k, &fi, &fo);
        if (d.conv_type == GNNB_CONV_GCN) {  // [bias, lin_weight]
            pend.push_back({img.size(), fi, fo});
            build_image(m->params[idx + 1].host.data(), fo, fo, fi, fi, 0, img);
            idx += 2;
        } else if (d.conv_type == GNNB_CONV_GIN) {  // [w0, b0, w1, b1]
            pend.push_back({img.size(), fi, fo});
            build_image(m->params[idx].host.data(), fo, fo, fi, fi, 0, img);
            pend.push_back({img.size(), fo, fo});
            build_image(m->params[idx + 2].host.data(), fo, fo, fo, fo, 0, img);
            idx += 4;
        } else if (d.conv_type == GNNB_CONV_SAGE) { // [lin_l.weight, lin_l.bias, lin_r.weight]
            pend.push_back({img.size(), fi, fo});
            build_image(m->params[idx].host.data(), fo, fo, fi, fi, 0, img);
            pend.push_back({img.size(), fi, fo});
            build_image(m->params[idx + 2].host.data(), fo, fo, fi, fi, 0, img);
            idx += 3;
        } else {
            // PNA: [pre_w (fi x 2fi), pre_b, post_w (fo x 13fi), post_b, lin_w (fo x fo), lin_b]
            const float *pre_w = m->params[idx].host.data(), *pre_b = m->params[idx + 1].host.data();
            const float *post_w = m->params[idx + 2].host.data(), *post_b = m->params[idx + 3].host.data();
            const float *lin_w = m->params[idx + 4].host.data(), *lin_b = m->params[idx + 5].host.data();
            const int fiP = (fi + 15) / 16 * 16;
            PnaPending q{};
            q.fiP = fiP; q.foP = fo;
            auto add = [&](const float *W, int N, int n_valid, int K, int ld, int col0) {
                PnaPending::Lin l{img.size(), K, N, 0, false};
                build_image(W, N, n_valid, K, ld, col0, img);
                return l;
            };
            auto add_bias = [&](PnaPending::Lin &l, const float *b, int n, int n_pad) {
                l.bias = img.size();
                l.has_bias = true;
                img.resize(img.size() + n_pad, 0.0f);
                for (int i = 0; i < n; i++) img[l.bias + i] = b[i];
            };
            q.pa = add(pre_w, fiP, fi, fi, 2 * fi, fi);        // W_nbr  = W_pre[:, fi:2fi]
            q.pb = add(pre_w, fiP, fi, fi, 2 * fi, 0);         // W_self = W_pre[:, 0:fi]
            add_bias(q.pb, pre_b, fi, fiP);
            q.ps = add(post_w, fo, fo, fi, 13 * fi, 0);        // self block of post_nn
            q.ng = (fi + 31) / 32;
            for (int g = 0; g < q.ng; g++) {
                const int gw = std::min(32, (fi - 32 * g + 15) / 16 * 16);
                q.gw[g] = gw;
                const int Kg = 4 * gw;
                // K' = 64 h + 32 fh + 8 q + j  <->  statistic q of feature 32 g + 16 h + 8 fh + j
                // (lib:1857-1875: the 12F block is [identity | amplification | attenuation] x
                // [max | min | mean | std] x F)
                for (int sc = 0; sc < 3; sc++) {
                    std::vector<float> Wg((size_t)fo * Kg, 0.0f);
                    for (int o = 0; o < fo; o++)
                        for (int h = 0; h < gw / 16; h++)
                            for (int st = 0; st < 4; st++)
                                for (int fh = 0; fh < 2; fh++)
                                    for (int j = 0; j < 8; j++) {
                                        const int f = 32 * g + 16 * h + 8 * fh + j;
                                        if (f < fi)
                                            Wg[(size_t)o * Kg + 64 * h + 32 * fh + 8 * st + j] =
                                                post_w[(size_t)o * 13 * fi + fi + (size_t)(sc * 4 + st) * fi + f];
                                    }
                    PnaPending::Lin l = add(Wg.data(), fo, fo, Kg, Kg, 0);
                    (sc == 0 ? q.gid : sc == 1 ? q.gamp : q.gatt)[g] = l;
                }
            }
            q.pl = add(lin_w, fo, fo, fo, fo, 0);
            add_bias(q.pl, lin_b, fo, fo);
            q.b_post = img.size();
            img.resize(img.size() + fo, 0.0f);
            for (int i = 0; i < fo; i++) img[q.b_post + i] = post_b[i];
            pna_pend.push_back(q);
            idx += 6;
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
                build_image(W, h.N, out, h.K[c], in, 128 * c, img);
            }
            h.bias = img.size();
            img.resize(img.size() + h.N, 0.0f);
            for (int i = 0; i < out; i++) img[h.bias + i] = b[i];
            hpend.push_back(h);
            in = out;
        }
    }
    int copies = 8;
    if (const char *e = getenv("GNNB_TC_WEIGHT_COPIES")) copies = std::max(1, std::min(64, atoi(e)));
    // replica stride: the image size rounded up to 256 B plus an odd number of 256-byte lines so
    // that equal offsets of different replicas fall on different L2 slices
    const size_t img_bytes = img.size() * sizeof(float);
    const size_t copy_bytes = (img_bytes + 255) / 256 * 256 + 256 * 37;
    int rc = plan->images.ensure(copy_bytes * (size_t)copies);
    for (int c = 0; c < copies && rc == GNNB_OK; c++) {
        cudaError_t e = cudaMemcpy(static_cast<unsigned char *>(plan->images.ptr) + c * copy_bytes,
                                   img.data(), img_bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = cuda_fail(e, "upload weight images", __FILE__, __LINE__);
    }
    p.img_copy_bytes = copy_bytes;
    p.img_copies = copies;
    if (rc == GNNB_OK) rc = plan->flag.ensure(sizeof(int));
    if (rc == GNNB_OK) rc = plan->pending.ensure((size_t)kNumSMs * HEAD_G * PLD * sizeof(float));
    if (rc == GNNB_OK && getenv("GNNB_FUSED_TIMING") != nullptr) {
        rc = plan->timing.ensure(32 * sizeof(unsigned long long));
        if (rc == GNNB_OK) cudaMemset(plan->timing.ptr, 0, 32 * sizeof(unsigned long long));
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
            t.KA = (q.K + WATOM_K - 1) / WATOM_K;
            return t;
        };
        if (is_pna) {
            const PnaPending &q = pna_pend[k];
            auto mkp = [&](const PnaPending::Lin &l) {
                TLinear t;
                t.img = base + l.off; t.bias = l.has_bias ? base + l.bias : nullptr; t.K = l.K; t.N = l.N;
                t.KA = (l.K + WATOM_K - 1) / WATOM_K;
                return t;
            };
            PnaLayer &P = p.pna[k];
            P.pa = mkp(q.pa); P.pb = mkp(q.pb); P.ps = mkp(q.ps); P.pl = mkp(q.pl);
            for (int g = 0; g < q.ng; g++) {
                P.gid[g] = mkp(q.gid[g]); P.gamp[g] = mkp(q.gamp[g]); P.gatt[g] = mkp(q.gatt[g]);
                P.gw[g] = q.gw[g];
            }
            P.b_post = base + q.b_post;
            P.ng = q.ng; P.fiP = q.fiP; P.foP = q.foP;
            continue;
        }
        p.l0[k] = mk(pend[pi++], L.a.bias);
        if (d.conv_type == GNNB_CONV_GIN) p.l1[k] = mk(pend[pi++], L.b.bias);
        else if (d.conv_type == GNNB_CONV_SAGE) p.l1[k] = mk(pend[pi++], nullptr);
    }
    p.pna_delta = d.pna_delta;
    // first shared-memory region: the bf16 ADJ tile, or PNA's fp32 A_u rows [128][PNA_LDA]
    p.r0_bytes = is_pna ? (int)((TM * PNA_LDA * sizeof(float) + 1023) / 1024 * 1024) : tc::PLANE_BYTES;
    for (int j = 0; j < d.mlp_num_linear; j++) {
        const HeadPending &h = hpend[j];
        p.hchunks[j] = h.nch;
        p.head_n[j] = h.n_true;
        for (int c = 0; c < h.nch; c++) {
            TLinear t;
            t.img = base + h.off[c]; t.bias = base + h.bias; t.K = h.K[c]; t.N = h.N;
            t.KA = (h.K[c] + WATOM_K - 1) / WATOM_K;
            p.hl[j][c] = t;
        }
    }
    // weight ring: 4 x 16 KB, or for PNA with N <= 80 (10 KB units, ~130 of them per tile) 8 x 10 KB
    p.slot_log2 = 2; p.slot_stride = 16384;
    if (is_pna && m->emb_dim() <= 80 && d.hidden_dim <= 80) { p.slot_log2 = 3; p.slot_stride = 80 * tc::ROW_BYTES; }
    if (const char *e = getenv("GNNB_TC_RING_SLOTS_LOG2")) {   // tuning hook (2 or 3)
        const int v = atoi(e);
        if (v == 2 || (v == 3 && p.slot_stride <= 10240)) p.slot_log2 = v;
    }
    p.ring_bytes = (p.slot_stride << p.slot_log2);
    plan->smem_bytes = 1024 + (size_t)p.r0_bytes + (size_t)NPLANES * tc::PLANE_BYTES + p.ring_bytes + CNT_BYTES +
                       sizeof(Misc);
    if (plan->smem_bytes > 232448) { delete plan; return GNNB_OK; }   // (cannot happen with the limits above)
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
        plan->images.release(); plan->bounds.release(); plan->pack_tmp.release(); plan->flag.release();
        plan->pending.release(); plan->timing.release();
        delete plan;
        m->fused_tc = nullptr;
    }
}

bool fused_tc_supports(const gnnb_model *m, int max_nodes_in_batch, int max_edges_in_batch)
{
    (void)max_edges_in_batch;   // edges stream through the multiplicity counters: no capacity limit
    return m->fused_tc != nullptr && max_nodes_in_batch <= MAX_NODES_PER_GRAPH;
}

int fused_tc_run(gnnb_model *m, const float *x, const int32_t *coo, const int64_t *node_ptr,
                 const int64_t *edge_ptr, int n_graphs, int64_t total_nodes, int max_nodes,
                 float *out, cudaStream_t s, int *launches, bool reset_status)
{
    TcPlan *plan = m->fused_tc;
    GNNB_REQUIRE(plan != nullptr, "tensor-core fused kernel not available for this model");
    if (max_nodes < 1) max_nodes = 1;
    GNNB_REQUIRE(max_nodes <= MAX_NODES_PER_GRAPH, "graph too large for the fused kernel");
    // device-side greedy packing of the graphs into tiles (three tiny kernels, no host sync)
    const int n_chunks = (n_graphs + PACK_CHUNK - 1) / PACK_CHUNK;
    // consecutive tiles hold > 128 rows together (or 128 graphs): an upper bound for the list
    const size_t max_tiles = (size_t)(2 * (total_nodes / TM) + 2 * ((int64_t)n_graphs / TM) + n_chunks + 4);
    GNNB_REQUIRE(max_tiles < ((size_t)1 << 30), "too many tiles");
    GNNB_TRY(plan->bounds.ensure(sizeof(int32_t) * (max_tiles + 2)));
    GNNB_TRY(plan->pack_tmp.ensure(sizeof(int32_t) * ((size_t)n_chunks * PACK_CHUNK + 2 * (size_t)n_chunks + 2)));
    int32_t *starts_tmp = plan->pack_tmp.as<int32_t>();
    int32_t *counts = starts_tmp + (size_t)n_chunks * PACK_CHUNK;
    int32_t *offsets = counts + n_chunks;
    int32_t *n_tiles_dev = offsets + n_chunks;
    if (reset_status) GNNB_CUDA(cudaMemsetAsync(plan->flag.ptr, 0, sizeof(int), s));
    tc_pack_chunk_kernel<<<n_chunks, 256, 0, s>>>(node_ptr, n_graphs, starts_tmp, counts);
    tc_pack_scan_kernel<<<1, 1024, 0, s>>>(counts, n_chunks, offsets, n_tiles_dev);
    tc_pack_compact_kernel<<<n_chunks, 256, 0, s>>>(starts_tmp, counts, offsets, n_chunks, n_graphs,
                                                    plan->bounds.as<int32_t>());
    GNNB_CUDA(cudaGetLastError());
    TcParams p = plan->params;
    p.x = x; p.coo = coo; p.node_ptr = node_ptr; p.edge_ptr = edge_ptr; p.n_graphs = n_graphs;
    p.out = out; p.tile_bounds = plan->bounds.as<int32_t>(); p.n_tiles_ptr = n_tiles_dev;
    p.error_flag = plan->flag.as<int>();
    p.pending = plan->pending.as<float>();
    p.timing = plan->timing.as<unsigned long long>();
    const int grid = n_graphs < kNumSMs ? n_graphs : kNumSMs;
    fused_tc_kernel<<<grid, CTA_THREADS, plan->smem_bytes, s>>>(p);
    GNNB_CUDA(cudaGetLastError());
    if (launches) *launches += 4;
    return GNNB_OK;
}

int fused_tc_status(gnnb_model *m, int *status)
{
    *status = 0;
    TcPlan *plan = m->fused_tc;
    if (plan == nullptr) return GNNB_OK;
    GNNB_CUDA(cudaMemcpy(status, plan->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    if (plan->timing.ptr != nullptr) {
        unsigned long long t[32];
        GNNB_CUDA(cudaMemcpy(t, plan->timing.ptr, sizeof(t), cudaMemcpyDeviceToHost));
        GNNB_CUDA(cudaMemset(plan->timing.ptr, 0, sizeof(t)));
        if (t[19])
            fprintf(stderr, "[gnnb fused-tc issuer] %llu GEMMs: waiting for the A operand %.1f%%, for weight "
                            "units %.1f%% of %.3g cycles inside gemm_issue (%.0f per GEMM)\n",
                    t[19], 100.0 * (double)t[16] / (double)t[18], 100.0 * (double)t[17] / (double)t[18],
                    (double)t[18], (double)t[18] / (double)t[19]);
        // (the row passes include their waits for the MMAs of the phase they read)
        const char *names[16] = {"stage", "adj+planes", "agg->A pass", "hidden/self pass", "pool", "head",
                                 "-", "output pass", "pna:x->A", "pna:wait pre", "pna:A_u->smem",
                                 "pna:stats", "pna:wait MMA", "pna:store+handoff", "pna:combine", "-"};
        unsigned long long tot = 0;
        for (int i = 0; i < 16; i++) tot += t[i];
        fprintf(stderr, "[gnnb fused-tc phases]");
        for (int i = 0; i < 16; i++)
            if (t[i])
                fprintf(stderr, " %s %.1f%%", names[i], tot ? 100.0 * (double)t[i] / (double)tot : 0.0);
        fprintf(stderr, " (total %.3g cycles over all CTAs)\n", (double)tot);
#ifdef GNNB_TC_SUBTIMING
        unsigned long long sub[8];
        cudaMemcpyFromSymbol(sub, g_sub, sizeof(sub));
        unsigned long long zero[8] = {0};
        cudaMemcpyToSymbol(g_sub, zero, sizeof(zero));
        fprintf(stderr, "[gnnb fused-tc sub] tmem-ld %.1f%% row-work %.1f%% st-wait %.1f%% mma-done-wait %.1f%%\n",
                100.0 * sub[0] / tot, 100.0 * sub[1] / tot, 100.0 * sub[2] / tot, 100.0 * sub[3] / tot);
#endif
    }
    return GNNB_OK;
}

}  // namespace gnnb
