// CSR (by destination) neighbor aggregation -- the gather/reduce half of every conv of
// gnn_builder_lib.h:  gcn_conv_agg (lib:1213-1289), gin_conv_agg (1389-1437), sage_conv_agg
// (2161-2209), pna_conv_agg (1750-1834), lg/simple (2350-2549).
//
// Kernel (2) of the design: a (sub-)warp per destination row, lanes across the feature dimension
// with float4 (16 B) gathers of each neighbor row, so one neighbor row of F=128 floats is one
// fully coalesced 512 B warp request; neighbor loop unrolled x4 so four independent gathers are
// in flight per lane.  Rows are bucketed by in-degree: rows above `heavy_threshold` are skipped
// here and handled by the chunked heavy-row kernels below (a warp per 128-neighbor chunk, chunk
// sums added in a fixed order: deterministic).
//
// HBM roofline: algorithmic bytes per layer = E*(4F + 4 [nbr idx] + 4 [dinv/deg]) +
// N*(4F self + 4F_out write + 8 [offset, degree]).  Neighbors are visited in table order, so in
// STRICT mode (no FMA contraction, reference scaling formula) sums round exactly like the
// reference's sum_incremental accumulators.
#include "kernels.h"

namespace gnnb {

namespace {

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float *p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void load_rw(const float *p) { v[0] = *p; }   // data this kernel also writes
    __device__ __forceinline__ void store(float *p) const { p[0] = v[0]; }
};
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float *p)
    {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void load_rw(const float *p)
    {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float *p) const
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

// L2 eviction-priority policies for the gathers of a large graph (hub_bit mode): rows of hub
// sources are kept (evict_last), everything else streams through (evict_first).
__device__ __forceinline__ uint64_t l2_policy_keep()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
template <int VEC>
__device__ __forceinline__ void load_hint(Vec<VEC> &x, const float *p, uint64_t pol);
template <>
__device__ __forceinline__ void load_hint<1>(Vec<1> &x, const float *p, uint64_t pol)
{
    asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;\n" : "=f"(x.v[0]) : "l"(p), "l"(pol));
}
template <>
__device__ __forceinline__ void load_hint<4>(Vec<4> &x, const float *p, uint64_t pol)
{
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
        : "=f"(x.v[0]), "=f"(x.v[1]), "=f"(x.v[2]), "=f"(x.v[3])
        : "l"(p), "l"(pol));
}

template <int MODE, bool STRICT>
__device__ __forceinline__ float neighbor_scale(const AggArgs &a, int deg_v, float dinv_v, int u)
{
    if (MODE == AGG_GCN) {
        if (STRICT) {
            const float di = __fadd_rn(1.0f, (float)deg_v);
            const float dj = __fadd_rn(1.0f, (float)__ldg(a.in_deg + u));
            return __fdiv_rn(1.0f, __fsqrt_rn(__fmul_rn(di, dj)));  // lib:1249-1252
        }
        return __ldg(a.dinv + u);  // rsqrt(1+d_u); the rsqrt(1+d_v) factor is applied once at the end
    }
    if (MODE == AGG_LG) {
        return __fdiv_rn(1.0f, __fsqrt_rn((float)(deg_v * __ldg(a.in_deg + u))));  // lib:2386
    }
    return 1.0f;
}

template <int VEC, int MODE, bool STRICT>
__device__ __forceinline__ void accumulate(Vec<VEC> &acc, const Vec<VEC> &x, float scale)
{
#pragma unroll
    for (int i = 0; i < VEC; i++) {
        if (MODE == AGG_GCN || MODE == AGG_LG)
            acc.v[i] = mac<STRICT>(acc.v[i], x.v[i], scale);
        else
            acc.v[i] = __fadd_rn(acc.v[i], x.v[i]);
    }
}

// partial sum over neighbors [k0, k1) of row v for the feature chunk starting at column c
template <int VEC, int MODE, bool STRICT>
__device__ __forceinline__ void gather_range(const AggArgs &a, int off, int k0, int k1, int deg_v,
                                             float dinv_v, int c, Vec<VEC> &acc)
{
    int k = k0;
    if (!STRICT && a.hub_bit) {
        const uint64_t keep = l2_policy_keep(), stream = l2_policy_stream();
        for (; k + 4 <= k1; k += 4) {
            int u[4];
            Vec<VEC> x[4];
            float sc[4];
#pragma unroll
            for (int j = 0; j < 4; j++) u[j] = __ldg(a.nbr + off + k + j);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t pol = u[j] < 0 ? keep : stream;
                u[j] &= 0x7fffffff;
                load_hint<VEC>(x[j], a.x + (size_t)u[j] * a.ldx + c, pol);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) sc[j] = neighbor_scale<MODE, STRICT>(a, deg_v, dinv_v, u[j]);
#pragma unroll
            for (int j = 0; j < 4; j++) accumulate<VEC, MODE, STRICT>(acc, x[j], sc[j]);
        }
        for (; k < k1; k++) {
            int u = __ldg(a.nbr + off + k);
            const uint64_t pol = u < 0 ? keep : stream;
            u &= 0x7fffffff;
            Vec<VEC> x;
            load_hint<VEC>(x, a.x + (size_t)u * a.ldx + c, pol);
            accumulate<VEC, MODE, STRICT>(acc, x, neighbor_scale<MODE, STRICT>(a, deg_v, dinv_v, u));
        }
        return;
    }
    for (; k + 4 <= k1; k += 4) {
        int u[4];
        Vec<VEC> x[4];
        float sc[4];
#pragma unroll
        for (int j = 0; j < 4; j++) u[j] = __ldg(a.nbr + off + k + j);
#pragma unroll
        for (int j = 0; j < 4; j++) x[j].load(a.x + (size_t)u[j] * a.ldx + c);
#pragma unroll
        for (int j = 0; j < 4; j++) sc[j] = neighbor_scale<MODE, STRICT>(a, deg_v, dinv_v, u[j]);
#pragma unroll
        for (int j = 0; j < 4; j++) accumulate<VEC, MODE, STRICT>(acc, x[j], sc[j]);
    }
    for (; k < k1; k++) {
        const int u = __ldg(a.nbr + off + k);
        Vec<VEC> x;
        x.load(a.x + (size_t)u * a.ldx + c);
        accumulate<VEC, MODE, STRICT>(acc, x, neighbor_scale<MODE, STRICT>(a, deg_v, dinv_v, u));
    }
}

// self term / normalisation and store
template <int VEC, int MODE, bool STRICT>
__device__ __forceinline__ void finish_row(const AggArgs &a, int v, int deg_v, float dinv_v, int c,
                                           Vec<VEC> &acc)
{
    if (!STRICT && a.no_finish) {   // first part of a split CSR: the plain partial sum
        acc.store(a.out + (size_t)v * a.ldo + c);
        return;
    }
    if (MODE == AGG_GCN) {
        Vec<VEC> xs;
        xs.load(a.x + (size_t)(v + a.row_base) * a.ldx + c);
        if (STRICT) {
            const float di = __fadd_rn(1.0f, (float)deg_v);
            const float ss = __fdiv_rn(1.0f, __fsqrt_rn(__fmul_rn(di, di)));  // lib:1266-1267
#pragma unroll
            for (int i = 0; i < VEC; i++)
                acc.v[i] = __fadd_rn(acc.v[i], __fmul_rn(xs.v[i], ss));
        } else {
            const float ss = dinv_v * dinv_v;
#pragma unroll
            for (int i = 0; i < VEC; i++) acc.v[i] = fmaf(xs.v[i], ss, acc.v[i] * dinv_v);
        }
    } else if (MODE == AGG_GIN) {
        Vec<VEC> xs;
        xs.load(a.x + (size_t)(v + a.row_base) * a.ldx + c);
        const float s = __fadd_rn(1.0f, a.eps);  // lib:1522
#pragma unroll
        for (int i = 0; i < VEC; i++) acc.v[i] = __fadd_rn(acc.v[i], __fmul_rn(xs.v[i], s));
    } else if (MODE == AGG_MEAN) {
        if (deg_v > 0) {
            const float d = (float)deg_v;
#pragma unroll
            for (int i = 0; i < VEC; i++) acc.v[i] = __fdiv_rn(acc.v[i], d);  // lib:661
        }
    }
    acc.store(a.out + (size_t)v * a.ldo + c);
}

// LPR lanes cooperate on one destination row; 32/LPR rows per warp.
template <int VEC, int LPR, int MODE, bool STRICT>
__global__ void __launch_bounds__(256) agg_rows_kernel(const AggArgs a)
{
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int lg = lane % LPR;
    const int sub = lane / LPR;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int64_t warp_stride = (int64_t)gridDim.x * warps_per_block;
    for (int64_t row0 = warp_global * ROWS_PER_WARP; row0 < a.n; row0 += warp_stride * ROWS_PER_WARP) {
        const int v = (int)row0 + sub;
        if (v >= a.n) continue;
        const int deg_v = __ldg(a.in_deg + v);
        const int len_v = (!STRICT && a.counts != nullptr) ? __ldg(a.counts + v) : deg_v;
        if (a.n_heavy > 0 && len_v > a.heavy_threshold) continue;  // CTA-per-row kernel does these
        const int off = __ldg(a.offsets + v);
        const float dinv_v = (MODE == AGG_GCN && !STRICT) ? __ldg(a.dinv + v + a.row_base) : 0.0f;
        for (int c = lg * VEC; c < a.F; c += LPR * VEC) {
            Vec<VEC> acc;
            if (!STRICT && a.accumulate) {
                acc.load_rw(a.out + (size_t)v * a.ldo + c);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; i++) acc.v[i] = 0.0f;
            }
            gather_range<VEC, MODE, STRICT>(a, off, 0, len_v, deg_v, dinv_v, c, acc);
            finish_row<VEC, MODE, STRICT>(a, v, deg_v, dinv_v, c, acc);
        }
    }
}

// Heavy rows (longer than the threshold; on the 2M-node power-law graph 0.5 % of the rows hold
// 49 % of the edges): the neighbor list of row h is cut into chunks of kHeavyChunk neighbors and
// every chunk is one warp's work item (grid-stride over all chunks of all heavy rows, so a
// 300-neighbor row and a 100 000-neighbor hub load the machine alike); the chunk sums go to
// partial[chunk][F].  A second kernel adds a row's chunks in order and applies the self term /
// normalisation, so the result does not depend on scheduling.  (Round 1 cut every heavy row into
// a fixed number of slices x 8 warps: rows just above the threshold then gave each warp one or two
// neighbors, and the kernel reached 38 % of the DRAM rate the light-row kernel reaches.)
template <int VEC, int MODE>
__global__ void __launch_bounds__(256) agg_heavy_chunk_kernel(const AggArgs a)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int warp_stride = gridDim.x * warps_per_block;
    for (int chunk = warp_global; chunk < a.heavy_chunks; chunk += warp_stride) {
        const int h = __ldg(a.heavy_chunk_row + chunk);
        const int v = __ldg(a.heavy_rows + h);
        const int j = chunk - __ldg(a.heavy_chunk_base + h);
        const int deg_v = __ldg(a.in_deg + v);
        const int len_v = a.counts != nullptr ? __ldg(a.counts + v) : deg_v;
        const int off = __ldg(a.offsets + v);
        const float dinv_v = (MODE == AGG_GCN) ? __ldg(a.dinv + v + a.row_base) : 0.0f;
        const int k0 = j * kHeavyChunk, k1 = min(len_v, k0 + kHeavyChunk);
        float *dst = a.heavy_partial + (size_t)chunk * a.F;
        for (int c = lane * VEC; c < a.F; c += 32 * VEC) {
            Vec<VEC> acc;
#pragma unroll
            for (int i = 0; i < VEC; i++) acc.v[i] = 0.0f;
            gather_range<VEC, MODE, false>(a, off, k0, k1, deg_v, dinv_v, c, acc);
            acc.store(dst + c);
        }
    }
}

// One CTA per heavy row: its 8 warps add contiguous ranges of the row's chunk sums (a hub with
// 100 000 neighbors has 782 of them: one warp walking them alone took ~100 us, a constant that
// did not shrink with the GPU count), the 8 range sums are then added in warp order through shared
// memory and the row is finished -- a fixed order, so the result does not depend on scheduling.
template <int VEC, int MODE>
__global__ void __launch_bounds__(256) agg_heavy_combine_kernel(const AggArgs a)
{
    extern __shared__ float ranges[];   // [8][F]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int h = blockIdx.x; h < a.n_heavy; h += gridDim.x) {
        const int v = __ldg(a.heavy_rows + h);
        const int deg_v = __ldg(a.in_deg + v);
        const int len_v = a.counts != nullptr ? __ldg(a.counts + v) : deg_v;
        const int nch = (len_v + kHeavyChunk - 1) / kHeavyChunk;
        const float *src = a.heavy_partial + (size_t)__ldg(a.heavy_chunk_base + h) * a.F;
        const float dinv_v = (MODE == AGG_GCN) ? __ldg(a.dinv + v + a.row_base) : 0.0f;
        const int per = (nch + 7) / 8;
        const int j0 = min(nch, warp * per), j1 = min(nch, j0 + per);
        for (int c = lane * VEC; c < a.F; c += 32 * VEC) {
            Vec<VEC> acc;
#pragma unroll
            for (int i = 0; i < VEC; i++) acc.v[i] = 0.0f;
            int j = j0;
            for (; j + 4 <= j1; j += 4) {     // four independent loads in flight, added in order
                Vec<VEC> t[4];
#pragma unroll
                for (int q = 0; q < 4; q++) t[q].load_rw(src + (size_t)(j + q) * a.F + c);
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int i = 0; i < VEC; i++) acc.v[i] += t[q].v[i];
            }
            for (; j < j1; j++) {
                Vec<VEC> t;
                t.load_rw(src + (size_t)j * a.F + c);
#pragma unroll
                for (int i = 0; i < VEC; i++) acc.v[i] += t.v[i];
            }
#pragma unroll
            for (int i = 0; i < VEC; i++) ranges[warp * a.F + c + i] = acc.v[i];
        }
        __syncthreads();
        if (warp == 0) {
            for (int c = lane * VEC; c < a.F; c += 32 * VEC) {
                Vec<VEC> acc;
                if (a.accumulate) {
                    acc.load_rw(a.out + (size_t)v * a.ldo + c);
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; i++) acc.v[i] = 0.0f;
                }
#pragma unroll
                for (int w = 0; w < 8; w++)
#pragma unroll
                    for (int i = 0; i < VEC; i++) acc.v[i] += ranges[w * a.F + c + i];
                finish_row<VEC, MODE, false>(a, v, deg_v, dinv_v, c, acc);
            }
        }
        __syncthreads();
    }
}

template <int VEC, int LPR, bool STRICT>
int launch_mode(const AggArgs &a, int grid, cudaStream_t s)
{
    switch (a.mode) {
    case AGG_GCN: agg_rows_kernel<VEC, LPR, AGG_GCN, STRICT><<<grid, 256, 0, s>>>(a); break;
    case AGG_GIN: agg_rows_kernel<VEC, LPR, AGG_GIN, STRICT><<<grid, 256, 0, s>>>(a); break;
    case AGG_MEAN: agg_rows_kernel<VEC, LPR, AGG_MEAN, STRICT><<<grid, 256, 0, s>>>(a); break;
    case AGG_SUM: agg_rows_kernel<VEC, LPR, AGG_SUM, STRICT><<<grid, 256, 0, s>>>(a); break;
    case AGG_LG: agg_rows_kernel<VEC, LPR, AGG_LG, STRICT><<<grid, 256, 0, s>>>(a); break;
    default: set_error("unknown aggregation mode"); return GNNB_ERR_INVALID;
    }
    return GNNB_OK;
}

template <int VEC, bool STRICT>
int launch_lpr(const AggArgs &a, int lpr, int grid, cudaStream_t s)
{
    switch (lpr) {
    case 1: return launch_mode<VEC, 1, STRICT>(a, grid, s);
    case 2: return launch_mode<VEC, 2, STRICT>(a, grid, s);
    case 4: return launch_mode<VEC, 4, STRICT>(a, grid, s);
    case 8: return launch_mode<VEC, 8, STRICT>(a, grid, s);
    case 16: return launch_mode<VEC, 16, STRICT>(a, grid, s);
    default: return launch_mode<VEC, 32, STRICT>(a, grid, s);
    }
}

template <int VEC, int MODE>
void launch_heavy_mode(const AggArgs &a, cudaStream_t s)
{
    const int cap = a.short_ctas ? (1 << 30) : kNumSMs * 8;   // 8 resident CTAs per SM, grid-stride beyond
    const int g1 = min((a.heavy_chunks + 7) / 8, cap), g2 = min(a.n_heavy, cap);
    agg_heavy_chunk_kernel<VEC, MODE><<<g1, 256, 0, s>>>(a);
    agg_heavy_combine_kernel<VEC, MODE><<<g2, 256, sizeof(float) * 8 * (size_t)a.F, s>>>(a);
}

template <int VEC>
int launch_heavy(const AggArgs &a, cudaStream_t s)
{
    switch (a.mode) {
    case AGG_GCN: launch_heavy_mode<VEC, AGG_GCN>(a, s); break;
    case AGG_GIN: launch_heavy_mode<VEC, AGG_GIN>(a, s); break;
    case AGG_MEAN: launch_heavy_mode<VEC, AGG_MEAN>(a, s); break;
    case AGG_SUM: launch_heavy_mode<VEC, AGG_SUM>(a, s); break;
    default: set_error("heavy-row path: unsupported mode"); return GNNB_ERR_INVALID;
    }
    return GNNB_OK;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int launch_agg(const AggArgs &a_in, bool strict, cudaStream_t s, int *launches)
{
    AggArgs a = a_in;
    if (a.n <= 0) return GNNB_OK;
    GNNB_REQUIRE(a.F > 0, "aggregation: feature size must be positive");
    if (strict) {
        GNNB_REQUIRE(a.counts == nullptr && !a.accumulate && !a.no_finish && !a.hub_bit,
                     "split-CSR / hub-hint aggregation is FAST-mode only");
        a.n_heavy = 0;
    }
    if (a.mode == AGG_GCN && !strict)
        GNNB_REQUIRE(a.dinv != nullptr, "gcn aggregation needs the dinv table in fast mode");
    const bool v4 = (a.F % 4 == 0) && (a.ldx % 4 == 0) && (a.ldo % 4 == 0) && aligned16(a.x) &&
                    aligned16(a.out);
    const int vec = v4 ? 4 : 1;
    int lpr = 1;
    while (lpr < 32 && lpr * vec < a.F) lpr *= 2;
    const int rows_per_block = 8 * (32 / lpr);
    int64_t grid64 = ceil_div64(a.n, rows_per_block);
    const int64_t cap = a.short_ctas ? (int64_t)1 << 30
                                     : (int64_t)kNumSMs * 8 * 4;  // 8 resident CTAs/SM x 4 waves, grid-stride
    int grid = (int)(grid64 < cap ? grid64 : cap);
    int rc;
    if (vec == 4)
        rc = strict ? launch_lpr<4, true>(a, lpr, grid, s) : launch_lpr<4, false>(a, lpr, grid, s);
    else
        rc = strict ? launch_lpr<1, true>(a, lpr, grid, s) : launch_lpr<1, false>(a, lpr, grid, s);
    GNNB_TRY(rc);
    GNNB_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    if (a.n_heavy > 0) {
        GNNB_REQUIRE(a.heavy_rows != nullptr && a.heavy_partial != nullptr && a.heavy_chunks > 0 &&
                         a.heavy_chunk_base != nullptr && a.heavy_chunk_row != nullptr,
                     "heavy row list / partial buffer missing");
        GNNB_TRY(vec == 4 ? launch_heavy<4>(a, s) : launch_heavy<1>(a, s));
        GNNB_CUDA(cudaGetLastError());
        if (launches) *launches += 2;
    }
    return GNNB_OK;
}

// ------------------------------------------------------------------------------------ PNA
//
// lib:1750-1834 with the legal rewrite W_pre.[x_v || x_u] + b = (W_self.x_v + b) + W_nbr.x_u
// (SURVEY section 7): `ab` holds A = X.W_nbr^T in columns [0,F) and B = X.W_self^T + b_pre in
// columns [F,2F), so the per-edge transformed message is t = A_u + B_v and the aggregation is a
// pure gather-reduce: max, min, mean and Welford variance (lib:677-705, population variance,
// std = sqrt(var + 1e-5); 0/0 = NaN for in-degree 0 exactly like the reference float build).
// The three degree scalers (lib:1973-1984, 2081-2089) are applied here and the 12F concat
// written in the reference's order (lib:1857-1875 minus the leading self block).
namespace {

template <int VEC, int LPR>
__global__ void __launch_bounds__(256) pna_agg_kernel(const PnaAggArgs a)
{
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int lg = lane % LPR, sub = lane / LPR;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int64_t warp_stride = (int64_t)gridDim.x * warps_per_block;
    const int F = a.F, ld = 2 * a.F;
    for (int64_t row0 = warp_global * ROWS_PER_WARP; row0 < a.n; row0 += warp_stride * ROWS_PER_WARP) {
        const int v = (int)row0 + sub;
        if (v >= a.n) continue;
        const int deg = __ldg(a.in_deg + v);
        const int off = __ldg(a.offsets + v);
        const int clamped = deg < 1 ? 1 : deg;
        const float lg1 = logf((float)(clamped + 1));
        const float amp = __fdiv_rn(lg1, a.delta), att = __fdiv_rn(a.delta, lg1);
        for (int c = lg * VEC; c < F; c += LPR * VEC) {
            Vec<VEC> b, vmax, vmin, vsum, wmean, wm2;
            b.load(a.ab + (size_t)v * ld + F + c);
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                vmax.v[i] = 0.0f; vmin.v[i] = 0.0f; vsum.v[i] = 0.0f;
                wmean.v[i] = 0.0f; wm2.v[i] = 0.0f;
            }
            for (int k = 0; k < deg; k++) {
                const int u = __ldg(a.nbr + off + k);
                Vec<VEC> t;
                t.load(a.ab + (size_t)u * ld + c);
                const float cnt = (float)(k + 1);
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    const float tv = __fadd_rn(t.v[i], b.v[i]);
                    vmax.v[i] = (k == 0 || tv > vmax.v[i]) ? tv : vmax.v[i];  // lib:748-759
                    vmin.v[i] = (k == 0 || tv < vmin.v[i]) ? tv : vmin.v[i];  // lib:784-795
                    vsum.v[i] = __fadd_rn(vsum.v[i], tv);
                    const float d = __fsub_rn(tv, wmean.v[i]);                    // lib:693
                    wmean.v[i] = __fadd_rn(wmean.v[i], __fdiv_rn(d, cnt));        // lib:694
                    wm2.v[i] = __fadd_rn(wm2.v[i], __fmul_rn(d, __fsub_rn(tv, wmean.v[i])));
                }
            }
            Vec<VEC> o[12];
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                const float mean = deg > 0 ? __fdiv_rn(vsum.v[i], (float)deg) : 0.0f;
                const float var = __fdiv_rn(wm2.v[i], (float)deg);                 // lib:702
                const float sd = __fsqrt_rn(__fadd_rn(var, 1e-5f));                // lib:703
                const float q[4] = {vmax.v[i], vmin.v[i], mean, sd};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    o[j].v[i] = q[j];
                    o[4 + j].v[i] = __fmul_rn(amp, q[j]);
                    o[8 + j].v[i] = __fmul_rn(att, q[j]);
                }
            }
            if (a.compact) {
                float *dst = a.cat12 + (size_t)v * 4 * F + c;
#pragma unroll
                for (int j = 0; j < 4; j++) o[j].store(dst + (size_t)j * F);
            } else {
                float *dst = a.cat12 + (size_t)v * 12 * F + c;
#pragma unroll
                for (int j = 0; j < 12; j++) o[j].store(dst + (size_t)j * F);
            }
        }
    }
}

template <int VEC>
int launch_pna_lpr(const PnaAggArgs &a, int lpr, int grid, cudaStream_t s)
{
    switch (lpr) {
    case 1: pna_agg_kernel<VEC, 1><<<grid, 256, 0, s>>>(a); break;
    case 2: pna_agg_kernel<VEC, 2><<<grid, 256, 0, s>>>(a); break;
    case 4: pna_agg_kernel<VEC, 4><<<grid, 256, 0, s>>>(a); break;
    case 8: pna_agg_kernel<VEC, 8><<<grid, 256, 0, s>>>(a); break;
    case 16: pna_agg_kernel<VEC, 16><<<grid, 256, 0, s>>>(a); break;
    default: pna_agg_kernel<VEC, 32><<<grid, 256, 0, s>>>(a); break;
    }
    return GNNB_OK;
}

}  // namespace

int launch_pna_agg(const PnaAggArgs &a, cudaStream_t s, int *launches)
{
    if (a.n <= 0) return GNNB_OK;
    const bool v4 = (a.F % 4 == 0) && aligned16(a.ab) && aligned16(a.cat12);
    const int vec = v4 ? 4 : 1;
    int lpr = 1;
    while (lpr < 32 && lpr * vec < a.F) lpr *= 2;
    const int rows_per_block = 8 * (32 / lpr);
    int64_t grid64 = ceil_div64(a.n, rows_per_block);
    const int64_t cap = (int64_t)kNumSMs * 8 * 4;
    const int grid = (int)(grid64 < cap ? grid64 : cap);
    GNNB_TRY(vec == 4 ? launch_pna_lpr<4>(a, lpr, grid, s) : launch_pna_lpr<1>(a, lpr, grid, s));
    GNNB_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    return GNNB_OK;
}

// ------------------------------------------------------------------------------------ GINE
//
// lib:1555-1623 gine_conv_agg: agg_v = sum_k relu(x[u_k] + proj[eid_k]) over the in-edges of v in
// table order, where proj = edge_feat . W_e^T + b_e was computed per edge by one GEMM; then the GIN
// self term (1 + eps) x_v (lib:1519-1529).  A (sub-)warp per destination row like agg_rows_kernel;
// additions are separately rounded in reference order, so STRICT and FAST agree bit for bit.
namespace {

template <int VEC, int LPR>
__global__ void __launch_bounds__(256) gine_agg_kernel(const GineAggArgs a)
{
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int lg = lane % LPR, sub = lane / LPR;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int64_t warp_stride = (int64_t)gridDim.x * warps_per_block;
    const float self_scale = __fadd_rn(1.0f, a.eps);   // lib:1522
    for (int64_t row0 = warp_global * ROWS_PER_WARP; row0 < a.n; row0 += warp_stride * ROWS_PER_WARP) {
        const int v = (int)row0 + sub;
        if (v >= a.n) continue;
        const int deg = __ldg(a.in_deg + v);
        const int off = __ldg(a.offsets + v);
        for (int c = lg * VEC; c < a.F; c += LPR * VEC) {
            Vec<VEC> acc;
#pragma unroll
            for (int i = 0; i < VEC; i++) acc.v[i] = 0.0f;
            for (int k = 0; k < deg; k++) {
                const int u = __ldg(a.nbr + off + k);
                const int eid = __ldg(a.edge_index + off + k);
                Vec<VEC> xu, pe;
                xu.load(a.x + (size_t)u * a.ldx + c);
                pe.load(a.proj + (size_t)eid * a.ldp + c);
#pragma unroll
                for (int i = 0; i < VEC; i++) {
                    const float m = __fadd_rn(xu.v[i], pe.v[i]);                  // lib:1603
                    acc.v[i] = __fadd_rn(acc.v[i], m > 0.0f ? m : 0.0f);         // lib:1606-1610
                }
            }
            Vec<VEC> xs;
            xs.load(a.x + (size_t)v * a.ldx + c);
#pragma unroll
            for (int i = 0; i < VEC; i++) acc.v[i] = __fadd_rn(acc.v[i], __fmul_rn(xs.v[i], self_scale));
            acc.store(a.out + (size_t)v * a.ldo + c);
        }
    }
}

template <int VEC>
int launch_gine_lpr(const GineAggArgs &a, int lpr, int grid, cudaStream_t s)
{
    switch (lpr) {
    case 1: gine_agg_kernel<VEC, 1><<<grid, 256, 0, s>>>(a); break;
    case 2: gine_agg_kernel<VEC, 2><<<grid, 256, 0, s>>>(a); break;
    case 4: gine_agg_kernel<VEC, 4><<<grid, 256, 0, s>>>(a); break;
    case 8: gine_agg_kernel<VEC, 8><<<grid, 256, 0, s>>>(a); break;
    case 16: gine_agg_kernel<VEC, 16><<<grid, 256, 0, s>>>(a); break;
    default: gine_agg_kernel<VEC, 32><<<grid, 256, 0, s>>>(a); break;
    }
    return GNNB_OK;
}

}  // namespace

int launch_gine_agg(const GineAggArgs &a, cudaStream_t s, int *launches)
{
    if (a.n <= 0) return GNNB_OK;
    const bool v4 = (a.F % 4 == 0) && (a.ldx % 4 == 0) && (a.ldp % 4 == 0) && (a.ldo % 4 == 0) &&
                    aligned16(a.x) && aligned16(a.proj) && aligned16(a.out);
    const int vec = v4 ? 4 : 1;
    int lpr = 1;
    while (lpr < 32 && lpr * vec < a.F) lpr *= 2;
    const int rows_per_block = 8 * (32 / lpr);
    int64_t grid64 = ceil_div64(a.n, rows_per_block);
    const int64_t cap = (int64_t)kNumSMs * 8 * 4;
    const int grid = (int)(grid64 < cap ? grid64 : cap);
    GNNB_TRY(vec == 4 ? launch_gine_lpr<4>(a, lpr, grid, s) : launch_gine_lpr<1>(a, lpr, grid, s));
    GNNB_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    return GNNB_OK;
}

}  // namespace gnnb
