// Kernel (1): persistent whole-model fused kernel for molecular-sized graphs.
//
// Mirrors the FPGA dataflow of the generated top (model.cpp.jinja:686-766) on one SM: a CTA
// owns a *tile of graphs* whose node rows pack into <= 128 rows, stages their node features and
// COO lists into shared memory once, builds the degree / neighbor tables there (lib:1051-1124,
// stable order), and then runs every conv layer (aggregate in SMEM -> node-transform GEMM with
// the weights streamed through a cp.async double buffer -> bias / skip / activation in the
// epilogue), the global pooling and the MLP head without touching HBM again.  Only out_dim
// floats per graph are written.  Algorithmic HBM bytes per graph: 4 n F_in + 8 e + 8 read,
// 4 out_dim written; the kernel is bound by the fp32 FMA pipe, not by HBM.
//
// Tiling is computed on the device: tile t owns the graphs whose first node row lies in
// [t*window, (t+1)*window) with window = 128 - max_nodes + 1, so a tile never exceeds 128 rows
// and no host-side packing pass is needed (tile_bounds_kernel = one binary search per tile).
// Every row's result depends only on its own graph, so outputs are bit-identical however the
// batch is composed.
//
// Layout: X[128][132] (layer input / skip source / layer output), WK[128][132] (aggregate ->
// hidden, in place), both fp32 row-major with the row stride padded to 132 floats so that the
// float4 reads of 8 rows at one k hit distinct banks; W tiles [16][BN] double buffered.
#include "model.h"

#include <algorithm>
#include <cstdlib>

namespace gnnb {

namespace {

constexpr int TM = 128;        // node rows per tile
constexpr int LDX = 132;       // feature buffer row stride (floats)
constexpr int ECAP = 2048;     // edges per tile
constexpr int BK = 32;        // k rows per staged weight tile (2 barriers per tile)
constexpr int MAX_LAYERS = 8;
constexpr int MAX_HEAD = 6;
constexpr int HEAD_G = 16;     // graphs per pooling/head chunk
constexpr int NTHREADS = 256;
constexpr int MAX_DIM = 128;
constexpr int MAX_NODES_PER_GRAPH = 64;

struct FLinear {
    const float *Wt;
    const float *bias;
    int in, out, ldw;
};

struct FusedParams {
    int conv_type, num_layers, in_dim, skip, gnn_act, num_pools, pools[4];
    int mlp_num_linear, mlp_act, out_act, emb, mlp_out;
    float gin_eps;
    int fi[MAX_LAYERS], fo[MAX_LAYERS];
    FLinear l0[MAX_LAYERS], l1[MAX_LAYERS];
    FLinear head[MAX_HEAD];
    const float *x;
    const int32_t *coo;
    const int64_t *node_ptr, *edge_ptr;
    int n_graphs;
    float *out;
    const int32_t *tile_bounds;  // [n_tiles + 1]
    int n_tiles;
    int *error_flag;
    unsigned long long *timing;  // optional per-phase cycle counters (GNNB_FUSED_TIMING=1)
};

struct Smem {
    float X[TM * LDX];
    float WK[TM * LDX];
    float WS[2][BK * MAX_DIM];
    float dinv[TM];
    int deg[TM];
    int off[TM + 1];
    int rowg[TM];
    int grow[TM + 2];
    int gedge[TM + 2];
    int scan_tmp[8];
    unsigned short edges[ECAP];  // (dst_row << 8) | src_row, tile-local rows
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

// first index g in [0, n) with ptr[g] >= v (n if none)
__device__ __forceinline__ int lower_bound64(const int64_t *__restrict__ ptr, int n, int64_t v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void tile_bounds_kernel(const int64_t *__restrict__ node_ptr, int n_graphs, int window,
                                   int n_tiles, int32_t *__restrict__ bounds)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    bounds[t] = (t == n_tiles) ? n_graphs : lower_bound64(node_ptr, n_graphs, (int64_t)t * window);
}

// ---------------------------------------------------------------------------------------
// C[128][N] = epilogue( A1[128][K1] . W1t (+ A2[128][K2] . W2t) + bias )
// A operands live in shared memory (row stride LDX), weights are streamed from global/L2 through
// the WS double buffer.  Every output element is accumulated by one thread in ascending k
// starting from the bias (the reference's order, lib:852-903).  dst may alias A1/A2/skip: all
// K-loop reads complete (barrier) before the first write.
template <int BN>
__device__ __noinline__ void tile_gemm(Smem &sm, const float *A1, int K1, const float *W1t,
                                          int ldw1, const float *A2, int K2, const float *W2t,
                                          int ldw2, const float *bias, int N, const float *skip,
                                          int act, float *dst)
{
    constexpr int TN = BN / 16;
    constexpr int CW = TN >= 4 ? 4 : TN;
    constexpr int NG = TN / CW;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    float acc[8][TN];
#pragma unroll
    for (int j = 0; j < TN; j++) {
        const int col = (j / CW) * (BN / NG) + tx * CW + (j % CW);
        const float b = (bias != nullptr && col < N) ? __ldg(bias + col) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i][j] = b;
    }

    const int nt1 = (K1 + BK - 1) / BK;
    const int nt2 = (A2 != nullptr) ? (K2 + BK - 1) / BK : 0;
    const int nt = nt1 + nt2;

    auto issue = [&](int t, int buf) {
        const bool second = t >= nt1;
        const float *Wt = second ? W2t : W1t;
        const int ldw = second ? ldw2 : ldw1;
        const int K = second ? K2 : K1;
        const int k0 = (second ? t - nt1 : t) * BK;
        constexpr int GPR = BN / 4;  // 16-byte granules per k row
        float *ws = sm.WS[buf];
        for (int g = tid; g < BK * GPR; g += NTHREADS) {
            const int kk = g / GPR, c4 = (g % GPR) * 4;
            const bool valid = (k0 + kk < K) && (c4 < ldw);
            const float *src = valid ? Wt + (size_t)(k0 + kk) * ldw + c4 : Wt;
            cp_async16(ws + kk * BN + c4, src, valid);
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (int t = 0; t < nt; t++) {
        if (t + 1 < nt) {
            issue(t + 1, (t + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const bool second = t >= nt1;
        const float *A = second ? A2 : A1;
        const int k0 = (second ? t - nt1 : t) * BK;
        const float *ws = sm.WS[t & 1];
#pragma unroll 2
        for (int k4 = 0; k4 < BK; k4 += 4) {
            float a[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = (i / 4) * 64 + ty * 4 + (i % 4);
                const float4 v = *reinterpret_cast<const float4 *>(A + row * LDX + k0 + k4);
                a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                float b[TN];
#pragma unroll
                for (int q = 0; q < NG; q++) {
                    const float *wp = ws + (k4 + kk) * BN + q * (BN / NG) + tx * CW;
                    if constexpr (CW == 4) {
                        const float4 w = *reinterpret_cast<const float4 *>(wp);
                        b[q * 4 + 0] = w.x; b[q * 4 + 1] = w.y; b[q * 4 + 2] = w.z; b[q * 4 + 3] = w.w;
                    } else {
                        const float2 w = *reinterpret_cast<const float2 *>(wp);
                        b[0] = w.x; b[1] = w.y;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i][kk], b[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

    // epilogue: (+ skip) -> activation -> dst; columns [N, round_up(N,32)) are zeroed so that
    // the next layer's K tiles read zeros, never stale data
    const int npad = (N + BK - 1) / BK * BK;
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        const int row = (i / 4) * 64 + ty * 4 + (i % 4);
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int col = (j / CW) * (BN / NG) + tx * CW + (j % CW);
            float v = 0.0f;
#pragma unroll
            for (int ii = 0; ii < 8; ii++)
                if (ii == i) v = acc[ii][j];
            if (col < N) {
                if (skip != nullptr) v += skip[row * LDX + col];
                dst[row * LDX + col] = act_apply_compact(act, v);
            } else if (col < npad && col < LDX) {
                dst[row * LDX + col] = 0.0f;
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void gemm_dispatch(Smem &sm, const float *A1, int K1, const float *W1t,
                                              int ldw1, const float *A2, int K2, const float *W2t,
                                              int ldw2, const float *bias, int N,
                                              const float *skip, int act, float *dst)
{
    if (N > 64)
        tile_gemm<128>(sm, A1, K1, W1t, ldw1, A2, K2, W2t, ldw2, bias, N, skip, act, dst);
    else if (N > 32)
        tile_gemm<64>(sm, A1, K1, W1t, ldw1, A2, K2, W2t, ldw2, bias, N, skip, act, dst);
    else
        tile_gemm<32>(sm, A1, K1, W1t, ldw1, A2, K2, W2t, ldw2, bias, N, skip, act, dst);
}

// one MLP-head linear for up to HEAD_G (=16) graphs: thread = (graph tid/16, column groups
// cg*4 and 64+cg*4).  The weights are streamed through the same cp.async double buffer as the
// conv GEMMs (a first version read them with __ldg from L2 inside the k loop and was bound by
// that latency: ~300 cycles x K per linear).
__device__ __noinline__ void head_linear(Smem &sm, const float *A, int lda, int K, const float *Wt,
                                         int ldw, const float *bias, int N, int act,
                                         float *dst_smem, int ldo, float *dst_global, int ldg,
                                         int n_rows)
{
    constexpr int BN = MAX_DIM;
    const int tid = threadIdx.x;
    const int gi = tid >> 4, cg = tid & 15;
    float acc[2][4];
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = q * 64 + cg * 4 + j;
            acc[q][j] = (bias != nullptr && col < N) ? __ldg(bias + col) : 0.0f;
        }
    const int nt = (K + BK - 1) / BK;
    auto issue = [&](int t, int buf) {
        constexpr int GPR = BN / 4;
        float *ws = sm.WS[buf];
        const int k0 = t * BK;
        for (int g = tid; g < BK * GPR; g += NTHREADS) {
            const int kk = g / GPR, c4 = (g % GPR) * 4;
            const bool valid = (k0 + kk < K) && (c4 < ldw);
            const float *src = valid ? Wt + (size_t)(k0 + kk) * ldw + c4 : Wt;
            cp_async16(ws + kk * BN + c4, src, valid);
        }
        cp_async_commit();
    };
    const float *a = A + (gi < n_rows ? gi : 0) * lda;
    issue(0, 0);
#pragma unroll 1
    for (int t = 0; t < nt; t++) {
        if (t + 1 < nt) {
            issue(t + 1, (t + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float *ws = sm.WS[t & 1];
        const int k0 = t * BK;
        const int kn = min(BK, K - k0);
#pragma unroll 4
        for (int kk = 0; kk < kn; kk++) {
            const float av = a[k0 + kk];
            const float4 w0 = *reinterpret_cast<const float4 *>(ws + kk * BN + cg * 4);
            const float4 w1 = *reinterpret_cast<const float4 *>(ws + kk * BN + 64 + cg * 4);
            acc[0][0] = fmaf(av, w0.x, acc[0][0]); acc[0][1] = fmaf(av, w0.y, acc[0][1]);
            acc[0][2] = fmaf(av, w0.z, acc[0][2]); acc[0][3] = fmaf(av, w0.w, acc[0][3]);
            acc[1][0] = fmaf(av, w1.x, acc[1][0]); acc[1][1] = fmaf(av, w1.y, acc[1][1]);
            acc[1][2] = fmaf(av, w1.z, acc[1][2]); acc[1][3] = fmaf(av, w1.w, acc[1][3]);
        }
        __syncthreads();
    }
    if (gi < n_rows) {
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int col = q * 64 + cg * 4 + j;
                if (col >= N) continue;
                const float v = act_apply_compact(act, acc[q][j]);
                if (dst_global != nullptr) dst_global[(size_t)gi * ldg + col] = v;
                else dst_smem[gi * ldo + col] = v;
            }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS, 1) fused_model_kernel(const FusedParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_prev = clock64();
#define GNNB_PHASE(idx)                                                             \
    if (p.timing != nullptr && tid == 0) {                                          \
        const long long t_now = clock64();                                          \
        atomicAdd(p.timing + (idx), (unsigned long long)(t_now - t_prev));          \
        t_prev = t_now;                                                             \
    }

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        // ------------------------------------------------------------------ tile geometry
        const int g0 = __ldg(p.tile_bounds + tile), g1 = __ldg(p.tile_bounds + tile + 1);
        const int ng = g1 - g0;
        if (ng <= 0) continue;  // uniform across the CTA
        const int64_t row0 = __ldg(p.node_ptr + g0);
        const int64_t e0 = __ldg(p.edge_ptr + g0);
        const int rows = (int)(__ldg(p.node_ptr + g1) - row0);
        const int ne = (int)(__ldg(p.edge_ptr + g1) - e0);
        if (rows > TM || ne > ECAP || ng > TM) {
            if (tid == 0) atomicExch(p.error_flag, 1);
            continue;
        }
        __syncthreads();  // previous tile's readers are done with the shared buffers
        for (int i = tid; i <= ng; i += NTHREADS) {
            sm.grow[i] = (int)(__ldg(p.node_ptr + g0 + i) - row0);
            sm.gedge[i] = (int)(__ldg(p.edge_ptr + g0 + i) - e0);
        }
        // node features, zero padded to a multiple of BK columns (K tiles of the first GEMM)
        {
            const int F = p.in_dim, Fp = (F + BK - 1) / BK * BK;
            const float *src = p.x + (size_t)row0 * F;
            for (int idx = tid; idx < rows * Fp; idx += NTHREADS) {
                const int r = idx / Fp, c = idx - r * Fp;
                sm.X[r * LDX + c] = (c < F) ? __ldg(src + (size_t)r * F + c) : 0.0f;
            }
        }
        __syncthreads();
        for (int i = tid; i < ng; i += NTHREADS)
            for (int r = sm.grow[i]; r < sm.grow[i + 1]; r++) sm.rowg[r] = i;
        // edges -> tile-local rows
        {
            const int2 *coo = reinterpret_cast<const int2 *>(p.coo) + e0;
            for (int j = tid; j < ne; j += NTHREADS) {
                const int2 sd = __ldg(coo + j);
                int lo = 0, hi = ng;  // graph of edge j: largest i with gedge[i] <= j
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (sm.gedge[mid] <= j) lo = mid; else hi = mid;
                }
                const int base = sm.grow[lo], n_i = sm.grow[lo + 1] - base;
                if ((unsigned)sd.x >= (unsigned)n_i || (unsigned)sd.y >= (unsigned)n_i) {
                    atomicExch(p.error_flag, 2);  // edge endpoint outside its graph
                    sm.edges[j] = 0xffff;
                } else {
                    sm.edges[j] = (unsigned short)(((sd.y + base) << 8) | (sd.x + base));
                }
            }
        }
        __syncthreads();
        GNNB_PHASE(0)
        // ------------------------------------------------------------------ tables (lib:1051-1124)
        int my_deg = 0;
        if (tid < TM) {
            if (tid < rows) {
                const int gi = sm.rowg[tid];
                for (int j = sm.gedge[gi]; j < sm.gedge[gi + 1]; j++)
                    my_deg += ((sm.edges[j] >> 8) == tid) ? 1 : 0;
            }
            sm.deg[tid] = my_deg;
            sm.dinv[tid] = 1.0f / sqrtf(1.0f + (float)my_deg);
            // exclusive scan of the in-degrees over the 128 rows (4 warps)
            int incl = my_deg;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) sm.scan_tmp[warp] = incl;
            sm.off[tid] = incl - my_deg;  // warp-local for now
        }
        __syncthreads();
        if (tid < TM) {
            int base = 0;
            for (int w = 0; w < warp; w++) base += sm.scan_tmp[w];
            const int o = sm.off[tid] + base;
            sm.off[tid] = o;
            if (tid < rows) {
                const int gi = sm.rowg[tid];
                int pos = o;
                for (int j = sm.gedge[gi]; j < sm.gedge[gi + 1]; j++) {
                    const unsigned short ed = sm.edges[j];
                    if ((ed >> 8) == tid) sm.nbr[pos++] = (unsigned char)(ed & 0xff);
                }
            }
        }
        __syncthreads();
        GNNB_PHASE(1)

        // ------------------------------------------------------------------ conv layers
        for (int l = 0; l < p.num_layers; l++) {
            const int fi = p.fi[l], fo = p.fo[l];
            const int kp = (fi + BK - 1) / BK * BK;
            const bool do_skip = p.skip && l != 0 && l != p.num_layers - 1;  // cpp:269-279
            // aggregate: warp per row, lanes across features (float4)
            for (int r = warp; r < rows; r += NTHREADS / 32) {
                const int d = sm.deg[r], o = sm.off[r];
                for (int c = lane * 4; c < kp; c += 128) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int k = 0; k < d; k++) {
                        const int u = sm.nbr[o + k];
                        const float4 v = *reinterpret_cast<const float4 *>(&sm.X[u * LDX + c]);
                        if (p.conv_type == GNNB_CONV_GCN) {
                            const float s = sm.dinv[u];
                            acc.x = fmaf(v.x, s, acc.x); acc.y = fmaf(v.y, s, acc.y);
                            acc.z = fmaf(v.z, s, acc.z); acc.w = fmaf(v.w, s, acc.w);
                        } else {
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                    }
                    const float4 xs = *reinterpret_cast<const float4 *>(&sm.X[r * LDX + c]);
                    if (p.conv_type == GNNB_CONV_GCN) {  // lib:1249-1278, factorised
                        const float dv = sm.dinv[r], ss = dv * dv;
                        acc.x = fmaf(xs.x, ss, acc.x * dv); acc.y = fmaf(xs.y, ss, acc.y * dv);
                        acc.z = fmaf(xs.z, ss, acc.z * dv); acc.w = fmaf(xs.w, ss, acc.w * dv);
                    } else if (p.conv_type == GNNB_CONV_GIN) {  // lib:1519-1529
                        const float s = 1.0f + p.gin_eps;
                        acc.x += xs.x * s; acc.y += xs.y * s; acc.z += xs.z * s; acc.w += xs.w * s;
                    } else if (d > 0) {  // SAGE mean, lib:656-662
                        const float dd = (float)d;
                        acc.x /= dd; acc.y /= dd; acc.z /= dd; acc.w /= dd;
                    }
                    *reinterpret_cast<float4 *>(&sm.WK[r * LDX + c]) = acc;
                }
            }
            __syncthreads();
            GNNB_PHASE(2)
            const float *skip = do_skip ? sm.X : nullptr;
            if (p.conv_type == GNNB_CONV_GCN) {
                gemm_dispatch(sm, sm.WK, fi, p.l0[l].Wt, p.l0[l].ldw, nullptr, 0, nullptr, 4,
                              p.l0[l].bias, fo, skip, p.gnn_act, sm.X);
            } else if (p.conv_type == GNNB_CONV_GIN) {
                gemm_dispatch(sm, sm.WK, fi, p.l0[l].Wt, p.l0[l].ldw, nullptr, 0, nullptr, 4,
                              p.l0[l].bias, fo, nullptr, GNNB_ACT_RELU, sm.WK);  // lib:1537-1541
                gemm_dispatch(sm, sm.WK, fo, p.l1[l].Wt, p.l1[l].ldw, nullptr, 0, nullptr, 4,
                              p.l1[l].bias, fo, skip, p.gnn_act, sm.X);           // lib:1542
            } else {  // SAGE: lin_l(mean) + lin_r(x), lib:2316-2332
                gemm_dispatch(sm, sm.WK, fi, p.l0[l].Wt, p.l0[l].ldw, sm.X, fi, p.l1[l].Wt,
                              p.l1[l].ldw, p.l0[l].bias, fo, skip, p.gnn_act, sm.X);
            }
            GNNB_PHASE(3)
        }

        // ------------------------------------------------------------------ pooling + MLP head
        const int emb = p.emb, head_in = emb * p.num_pools;
        const int ldp = ((head_in + 3) & ~3) + 4;
        float *pooled = sm.WK;
        float *hb0 = sm.WK + HEAD_G * ldp;
        float *hb1 = hb0 + HEAD_G * LDX;
        for (int gc0 = 0; gc0 < ng; gc0 += HEAD_G) {
            const int gcn = min(HEAD_G, ng - gc0);
            for (int gi = warp; gi < gcn; gi += NTHREADS / 32) {
                const int r0 = sm.grow[gc0 + gi], r1 = sm.grow[gc0 + gi + 1];
                for (int c = lane; c < emb; c += 32) {
                    float sum = 0.0f, mx = 0.0f;
                    for (int r = r0; r < r1; r++) {
                        const float v = sm.X[r * LDX + c];
                        sum += v;
                        mx = (r == r0 || v > mx) ? v : mx;  // lib:748-759
                    }
                    for (int q = 0; q < p.num_pools; q++) {
                        float v;
                        if (p.pools[q] == GNNB_POOL_ADD) v = sum;
                        else if (p.pools[q] == GNNB_POOL_MEAN) v = (r1 > r0) ? sum / (float)(r1 - r0) : 0.0f;
                        else v = mx;
                        pooled[gi * ldp + q * emb + c] = v;
                    }
                }
            }
            __syncthreads();
            GNNB_PHASE(4)
            const float *hin = pooled;
            int hld = ldp, hk = head_in;
            for (int j = 0; j < p.mlp_num_linear; j++) {
                const bool last = j == p.mlp_num_linear - 1;
                float *hout = (j & 1) ? hb1 : hb0;
                head_linear(sm, hin, hld, hk, p.head[j].Wt, p.head[j].ldw, p.head[j].bias,
                            p.head[j].out, last ? p.out_act : p.mlp_act, hout, LDX,
                            last ? p.out + (size_t)(g0 + gc0) * p.mlp_out : nullptr, p.mlp_out, gcn);
                hin = hout; hld = LDX; hk = p.head[j].out;
            }
            GNNB_PHASE(5)
        }
    }
#undef GNNB_PHASE
}

}  // namespace

struct FusedPlan {
    FusedParams params{};
    DeviceBuf bounds, flag, timing;
    size_t smem_bytes = 0;
};

static FLinear to_f(const PackedLinear &l)
{
    FLinear f;
    f.Wt = l.Wt; f.bias = l.bias; f.in = l.in; f.out = l.out; f.ldw = l.ldw;
    return f;
}

int fused_prepare(gnnb_model *m)
{
    m->fused = nullptr;
    const gnnb_model_desc &d = m->d;
    const bool conv_ok = d.conv_type == GNNB_CONV_GCN || d.conv_type == GNNB_CONV_GIN ||
                         d.conv_type == GNNB_CONV_SAGE;
    const int head_in = m->emb_dim() * d.num_pools;
    if (!conv_ok || d.num_layers > MAX_LAYERS || d.mlp_num_linear > MAX_HEAD ||
        d.in_dim > MAX_DIM || d.hidden_dim > MAX_DIM || d.out_dim > MAX_DIM ||
        d.mlp_hidden > MAX_DIM || d.mlp_out > MAX_DIM || head_in > 512)
        return GNNB_OK;  // not an error: the layerwise path handles it
    FusedPlan *plan = new FusedPlan();
    FusedParams &p = plan->params;
    p.conv_type = d.conv_type; p.num_layers = d.num_layers; p.in_dim = d.in_dim; p.skip = d.skip;
    p.gnn_act = d.gnn_act; p.num_pools = d.num_pools;
    for (int i = 0; i < 4; i++) p.pools[i] = d.pools[i];
    p.mlp_num_linear = d.mlp_num_linear; p.mlp_act = d.mlp_act; p.out_act = d.out_act;
    p.emb = m->emb_dim(); p.mlp_out = d.mlp_out; p.gin_eps = d.gin_eps;
    for (int k = 0; k < d.num_layers; k++) {
        p.fi[k] = m->layers[k].fi; p.fo[k] = m->layers[k].fo;
        p.l0[k] = to_f(m->layers[k].a);
        p.l1[k] = to_f(m->layers[k].b);
    }
    for (int j = 0; j < d.mlp_num_linear; j++) p.head[j] = to_f(m->head[j]);
    plan->smem_bytes = sizeof(Smem);
    cudaError_t e = cudaFuncSetAttribute(fused_model_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)plan->smem_bytes);
    if (e != cudaSuccess) {
        delete plan;
        return cuda_fail(e, "cudaFuncSetAttribute(fused_model_kernel)", __FILE__, __LINE__);
    }
    int rc = plan->flag.ensure(sizeof(int));
    if (rc == GNNB_OK && getenv("GNNB_FUSED_TIMING") != nullptr) {
        rc = plan->timing.ensure(16 * sizeof(unsigned long long));
        if (rc == GNNB_OK) cudaMemset(plan->timing.ptr, 0, 16 * sizeof(unsigned long long));
    }
    if (rc != GNNB_OK) { delete plan; return rc; }
    m->fused = plan;
    return GNNB_OK;
}

void fused_release(gnnb_model *m)
{
    if (m->fused) {
        m->fused->bounds.release();
        m->fused->flag.release();
        m->fused->timing.release();
        delete m->fused;
        m->fused = nullptr;
    }
}

int fused_tile_rows(const gnnb_model *) { return TM; }

bool fused_supports(const gnnb_model *m, int max_nodes_in_batch, int max_edges_in_batch)
{
    return m->fused != nullptr && max_nodes_in_batch <= MAX_NODES_PER_GRAPH &&
           max_edges_in_batch <= ECAP / 4;
}

int fused_run(gnnb_model *m, const float *x, const int32_t *coo, const int64_t *node_ptr,
              const int64_t *edge_ptr, int n_graphs, int64_t total_nodes, int max_nodes, float *out,
              cudaStream_t s, int *launches)
{
    FusedPlan *plan = m->fused;
    GNNB_REQUIRE(plan != nullptr, "fused kernel not available for this model");
    if (max_nodes < 1) max_nodes = 1;
    GNNB_REQUIRE(max_nodes <= MAX_NODES_PER_GRAPH, "graph too large for the fused kernel");
    const int window = TM - max_nodes + 1;
    const int64_t n_tiles64 = total_nodes / window + 1;
    GNNB_REQUIRE(n_tiles64 < (1ll << 30), "too many tiles");
    const int n_tiles = (int)n_tiles64;
    GNNB_TRY(plan->bounds.ensure(sizeof(int32_t) * ((size_t)n_tiles + 1)));
    GNNB_CUDA(cudaMemsetAsync(plan->flag.ptr, 0, sizeof(int), s));
    tile_bounds_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, s>>>(node_ptr, n_graphs, window, n_tiles,
                                                                 plan->bounds.as<int32_t>());
    GNNB_CUDA(cudaGetLastError());
    FusedParams p = plan->params;
    p.x = x; p.coo = coo; p.node_ptr = node_ptr; p.edge_ptr = edge_ptr; p.n_graphs = n_graphs;
    p.out = out; p.tile_bounds = plan->bounds.as<int32_t>(); p.n_tiles = n_tiles;
    p.error_flag = plan->flag.as<int>();
    p.timing = plan->timing.as<unsigned long long>();
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;  // persistent: one CTA per SM
    fused_model_kernel<<<grid, NTHREADS, plan->smem_bytes, s>>>(p);
    GNNB_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
    return GNNB_OK;
}

// after the stream has been synchronised: 0 = ok, 1 = a tile overflowed its capacity (the caller
// re-runs the batch layerwise), 2 = an edge endpoint was outside its graph
int fused_status(gnnb_model *m, int *status)
{
    *status = 0;
    if (m->fused == nullptr) return GNNB_OK;
    GNNB_CUDA(cudaMemcpy(status, m->fused->flag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
    if (m->fused->timing.ptr != nullptr) {  // developer aid: per-phase cycles summed over CTAs
        unsigned long long t[16];
        GNNB_CUDA(cudaMemcpy(t, m->fused->timing.ptr, sizeof(t), cudaMemcpyDeviceToHost));
        GNNB_CUDA(cudaMemset(m->fused->timing.ptr, 0, sizeof(t)));
        const char *names[6] = {"stage", "tables", "aggregate", "gemm", "pool", "head"};
        unsigned long long tot = 0;
        for (int i = 0; i < 6; i++) tot += t[i];
        fprintf(stderr, "[gnnb fused phases]");
        for (int i = 0; i < 6; i++)
            fprintf(stderr, " %s %.1f%%", names[i], tot ? 100.0 * (double)t[i] / (double)tot : 0.0);
        fprintf(stderr, " (total %.3g cycles over all CTAs)\n", (double)tot);
    }
    return GNNB_OK;
}

}  // namespace gnnb
