// Host-side launchers of the CUDA kernels (internal).
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace gnnb {

// ---------------------------------------------------------------- graph tables (tables.cu)
struct TableWorkspace {
    DeviceBuf keys_in, keys_out, vals_in, vals_out, cub_tmp, heavy_partial, counters;
    DeviceBuf hub_hist, hub_cnt;
    void release_all()
    {
        DeviceBuf *b[] = {&keys_in, &keys_out, &vals_in, &vals_out, &cub_tmp,
                          &heavy_partial, &counters, &hub_hist, &hub_cnt};
        for (DeviceBuf *x : b) x->release();
    }
};

// edge_list [E][2] local ids; node_ptr/edge_ptr (device int64, G+1 entries) translate them to
// ids in the batch union (pass nullptr for a single graph).  Writes in/out degree, offsets
// (n entries), neighbor_table (e entries) and, when non-null, edge_index_table.
// node_base/edge_base: values of node_ptr[0]/edge_ptr[0] when the pointers address a sub-range
// (chunk) of a larger batch whose x / edge_list pointers were advanced accordingly.
int build_tables(const int32_t *edge_list, const int64_t *node_ptr, const int64_t *edge_ptr,
                 int64_t node_base, int64_t edge_base, int n_graphs, int n, int e, int32_t *in_deg, int32_t *out_deg, int32_t *offsets,
                 int32_t *nbr, int32_t *edge_index, TableWorkspace &ws, cudaStream_t s,
                 int *launches, int *bad = nullptr);
// `bad` (device int, optional): set to 1 when an edge endpoint lies outside its graph; such edges
// are dropped from the degree counts so the tables stay memory-safe, and the caller reports
// GNNB_ERR_INVALID (the fused kernels report the same condition as status 2)
// degree tables only (lib:1051-1083)
int build_degree_tables(const int32_t *edge_list, int n, int e, int32_t *in_deg, int32_t *out_deg,
                        cudaStream_t s, int *launches, int *bad = nullptr);
// offsets + neighbor table from given in-degree (lib:1086-1166)
int build_neighbor_tables(const int32_t *edge_list, const int32_t *in_deg, int n, int e,
                          int32_t *offsets, int32_t *nbr, int32_t *edge_index, TableWorkspace &ws,
                          cudaStream_t s, int *launches, int *bad = nullptr);
// CSR slice of a 1D row partition: edge_list holds the in-edges of the owned destination rows
// [row_begin, row_begin + n_local) with GLOBAL node ids; neighbors keep their global ids
int build_partition_tables(const int32_t *edge_list, int row_begin, int n_local, int e,
                           int32_t *in_deg_local, int32_t *offsets_local, int32_t *nbr_global,
                           TableWorkspace &ws, cudaStream_t s, int *launches);
// Degree bucketing: rows longer than `threshold` ("heavy" rows; on the 2M-node power-law graph
// 0.5 % of the rows hold 49 % of the edges) are cut into chunks of kHeavyChunk neighbors, one warp
// per chunk, and a second kernel adds a row's chunk sums in order (deterministic).
constexpr int kHeavyChunk = 128;
struct HeavyList {
    DeviceBuf rows;        // [n_heavy] row ids
    DeviceBuf chunk_base;  // [n_heavy] first chunk of the row
    DeviceBuf chunk_row;   // [n_chunks] index into rows[] of the chunk's row
    int n_heavy = 0, n_chunks = 0;
    void release() { rows.release(); chunk_base.release(); chunk_row.release(); n_heavy = n_chunks = 0; }
};
// fills `hl` from the row lengths (two small kernels and one host synchronisation)
int find_heavy_rows(const int32_t *lengths, int n, int threshold, TableWorkspace &ws, HeavyList &hl,
                    cudaStream_t s, int *launches);
// degree bucketing parameters + partial-sum scratch for the heavy-row kernels
constexpr int kHeavyThreshold = 256;
// (tuning hook: GNNB_HEAVY_THRESHOLD overrides the bucket boundary)
inline int heavy_threshold()
{
    static const int v = [] {
        const char *e = getenv("GNNB_HEAVY_THRESHOLD");
        const int t = e ? atoi(e) : 0;
        return t > 0 ? t : kHeavyThreshold;
    }();
    return v;
}
// partial-sum scratch of the heavy rows: [n_chunks][F] floats (grow-only)
int heavy_setup(TableWorkspace &ws, int n_chunks, int F);
// hub sources: copy of the neighbor table with bit 31 set on the most-referenced sources whose
// feature rows fit budget_bytes (L2-resident set of the aggregation); tables.cu
int mark_hub_sources(const int32_t *nbr_in, int32_t *nbr_out, int e, const int32_t *ref_cnt,
                     int n_src, size_t row_bytes, size_t budget_bytes, TableWorkspace &ws,
                     int *n_hubs_host, cudaStream_t s, int *launches);
// L2 budget for hub rows in bytes (GNNB_HUB_L2_MB overrides; 0 disables the hints)
inline size_t hub_l2_budget()
{
    static const size_t v = [] {
        const char *e = getenv("GNNB_HUB_L2_MB");
        const long mb = e ? atol(e) : 0;   // off by default: measured, it cuts DRAM bytes by 10 %
                                           // and no time (profiles/r2_agg_c5.txt), and costs the
                                           // out-degree atomics + a marking pass per run
        return (size_t)(mb > 0 ? mb : 0) << 20;
    }();
    return v;
}
// dinv[i] = 1 / sqrt(1 + in_deg[i])
int compute_dinv(const int32_t *in_deg, float *dinv, int n, cudaStream_t s, int *launches);

// ---------------------------------------------------------------- aggregation (agg.cu)
enum AggMode { AGG_GCN = 0, AGG_GIN = 1, AGG_MEAN = 2, AGG_SUM = 3, AGG_LG = 4 };

struct AggArgs {
    int mode;
    const float *x;        // [n][ldx]
    int ldx;
    int F;
    float *out;            // [n][ldo]
    int ldo;
    const int32_t *offsets, *nbr, *in_deg;
    const float *dinv;     // AGG_GCN fast mode only
    int n;
    float eps;             // AGG_GIN
    const int32_t *heavy_rows;  // optional list of rows handled by the chunked heavy-row kernels
    int n_heavy;
    int heavy_threshold;
    float *heavy_partial;  // [heavy_chunks][F] chunk sums of the heavy rows
    const int32_t *heavy_chunk_base, *heavy_chunk_row;
    int heavy_chunks;
    void set_heavy(const HeavyList &hl, float *partial)
    {
        heavy_rows = hl.rows.as<int32_t>(); n_heavy = hl.n_heavy;
        heavy_chunk_base = hl.chunk_base.as<int32_t>(); heavy_chunk_row = hl.chunk_row.as<int32_t>();
        heavy_chunks = hl.n_chunks; heavy_partial = partial;
    }
    int row_base;          // row-partitioned graphs: global id of local row 0 (x / dinv are global,
                           // offsets / in_deg / out are local); 0 otherwise
    // ---- FAST-mode extensions (all 0 / null by default)
    const int32_t *counts; // row lengths when they differ from in_deg (one PART of a split CSR: the
                           // rows' owned-source or halo-source edges); in_deg still feeds the mean
    int accumulate;        // start every row sum from what `out` already holds (second CSR part)
    int no_finish;         // store the plain partial sum: no self term / normalisation (first part)
    int short_ctas;        // 1: one pass per CTA instead of a capped grid-stride grid, so that the CTAs
                           // of a higher-priority kernel (the halo pack on the exchange stream) are
                           // placed as these retire instead of waiting for the whole kernel
    int hub_bit;           // 1: bit 31 of a neighbor entry marks a hub source (high out-degree);
                           // its row is loaded with an L2 evict_last policy, every other row with
                           // evict_first, so the hub rows stay resident in the 126 MB L2
};
int launch_agg(const AggArgs &a, bool strict, cudaStream_t s, int *launches);

// GINE aggregation (lib:1555-1623): sum_k relu(x[nbr_k] + proj[edge_index_k]) + (1 + eps) x_v
struct GineAggArgs {
    const float *x;        // [n][ldx]
    int ldx, F;
    const float *proj;     // [E][ldp] projected edge features (edge id order)
    int ldp;
    float *out;            // [n][ldo]
    int ldo;
    const int32_t *offsets, *nbr, *edge_index, *in_deg;
    int n;
    float eps;
};
int launch_gine_agg(const GineAggArgs &a, cudaStream_t s, int *launches);

struct PnaAggArgs {
    const float *ab;       // [n][2F]: cols [0,F) = x.W_nbr^T, cols [F,2F) = x.W_self^T + b_pre
    int F;
    float *cat12;          // [n][12F]: (max,min,mean,std) x (identity, amplification, attenuation)
    const int32_t *offsets, *nbr, *in_deg;
    int n;
    float delta;
    int compact;           // 1: write only the [n][4F] statistics (max,min,mean,std); the two degree
                           // scalers are applied by the consumer (gemm_tc.cu expands them on the fly)
};
int launch_pna_agg(const PnaAggArgs &a, cudaStream_t s, int *launches);

// ---------------------------------------------------------------- GEMM (gemm.cu)
// C[M][N] = act( A1[M][K1].W1t[K1][N] (+ A2[M][K2].W2t[K2][N]) + bias (+ skip) )
// weights are pre-transposed: Wt[k][n], row stride ldw (multiple of 4, zero padded)
struct GemmArgs {
    const float *A1; int lda1; int K1; const float *W1t; int ldw1;
    const float *A2; int lda2; int K2; const float *W2t; int ldw2;
    const float *img1, *img2;   // optional tensor-core weight images of W1 / W2 (gemm_tc.cu); when
                                // present (and FAST math) the GEMM runs on tcgen05, else on the FMA pipe
    // PNA "expand" mode (tensor-core path only): A2 holds the [M][K2/3] statistics and stands for
    // [A2 | amp_v A2 | att_v A2] (lib:1857-1875), amp_v = log(max(deg_v,1)+1)/delta, att_v = 1/amp_v
    const int32_t *expand_deg; float expand_delta;
    int second_separate;        // STRICT only: A2.W2t is a separate bias-free sum added last (SAGE)
    const float *bias;          // [N] or null
    const float *skip; int ldskip;  // added before the activation, or null
    int act;
    float *C; int ldc;
    int M; int N;
};
int launch_gemm(const GemmArgs &g, bool strict, cudaStream_t s, int *launches);
// gemm_tc.cu: tcgen05 (3xTF32) version of the same contract, for large M
size_t gemm_tc_image_floats(int K, int N);
int gemm_tc_build_image(const float *Wt, int ldw, int K, int N, float *img, cudaStream_t s);
bool gemm_tc_supported(const GemmArgs &g);
int launch_gemm_tc(const GemmArgs &g, cudaStream_t s, int *launches);
// Wt[k][n] (ld = ldw) from W[n][k] (reference layout); pads columns n..ldw with zeros
int launch_transpose_weight(const float *W, float *Wt, int out_size, int in_size, int ldw,
                            cudaStream_t s, int *launches);

// ---------------------------------------------------------------- pooling / misc (pool.cu)
// pooled[g][p*F + f] for pools[p] over rows node_ptr[g]..node_ptr[g+1]
int launch_pool(const float *x, int ldx, int F, const int64_t *node_ptr, int64_t node_base,
                int n_graphs, int64_t total_nodes, const int *pools, int num_pools, float *pooled,
                DeviceBuf &tmp, cudaStream_t s, int *launches, bool single_pass = false);
int launch_activation(int act, const float *x, float *y, size_t n, cudaStream_t s, int *launches);

}  // namespace gnnb
