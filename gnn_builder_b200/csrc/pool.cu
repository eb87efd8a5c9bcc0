// Global pooling (gnn_builder_lib.h:2709-2803) over the graphs of a batch, and the element-wise
// activation entry point (lib:501-509).
//
// pooled[g][p*F + f] = pool_p over the node rows of graph g, concatenated in list order like the
// generated compute_global_graph_pooling (model.cpp.jinja:440-448).  One CTA per (graph, split):
// thread f walks the graph's node rows in order (so with one split the sum rounds exactly like
// the reference's sum_incremental loop); huge graphs are cut into `splits` row ranges whose
// partials are combined in range order by a second kernel.  mean = sum / n (the value the
// reference's running mean holds after its last update); max starts from the first row and is
// 0 for an empty graph (lib:736-760).
#include "kernels.h"

namespace gnnb {

namespace {

struct PoolArgs {
    const float *x; int ldx; int F;
    const int64_t *node_ptr; int64_t node_base; int n_graphs;
    int pools[4]; int num_pools;
    float *pooled;     // [G][num_pools*F]
    float *partial;    // [G][splits][2][F] (sum, max) when splits > 1
    int splits;
};

__device__ __forceinline__ void write_pools(const PoolArgs &a, int g, int f, float sum, float mx,
                                            int64_t n)
{
    float *dst = a.pooled + (size_t)g * a.num_pools * a.F;
    for (int p = 0; p < a.num_pools; p++) {
        float v;
        if (a.pools[p] == GNNB_POOL_ADD) v = sum;
        else if (a.pools[p] == GNNB_POOL_MEAN) v = n > 0 ? __fdiv_rn(sum, (float)n) : 0.0f;
        else v = n > 0 ? mx : 0.0f;
        dst[(size_t)p * a.F + f] = v;
    }
}

__global__ void __launch_bounds__(128) pool_kernel(const PoolArgs a)
{
    const int split = blockIdx.y;
    for (int g = blockIdx.x; g < a.n_graphs; g += gridDim.x) {
        const int64_t n0 = __ldg(a.node_ptr + g) - a.node_base;
        const int64_t n1 = __ldg(a.node_ptr + g + 1) - a.node_base;
        const int64_t n = n1 - n0;
        const int64_t per = (n + a.splits - 1) / a.splits;
        const int64_t r0 = n0 + split * per;
        const int64_t r1 = (r0 + per < n1) ? r0 + per : n1;
        for (int f = threadIdx.x; f < a.F; f += blockDim.x) {
            float sum = 0.0f, mx = split == 0 ? 0.0f : -INFINITY;
            // lib:748-759: the first sample of the GRAPH initialises the maximum (a NaN there
            // sticks), every later sample only replaces it when it compares greater (a NaN never
            // does).  Only range 0 holds the graph's first sample.
            bool first = split == 0;
            for (int64_t r = r0; r < r1; r++) {
                const float v = __ldg(a.x + (size_t)r * a.ldx + f);
                sum = __fadd_rn(sum, v);
                mx = (first || v > mx) ? v : mx;
                first = false;
            }
            if (a.splits == 1) {
                write_pools(a, g, f, sum, mx, n);
            } else {
                float *p = a.partial + ((size_t)g * a.splits + split) * 2 * a.F;
                p[f] = sum;
                p[a.F + f] = (split == 0 && r1 <= r0) ? -INFINITY : mx;
            }
        }
    }
}

// Partials of one graph are combined by a CTA of 8 groups x 128 feature lanes: group q walks the
// splits q, q + 8, ... (independent loads, short serial chain), the 8 group results are then added
// in group order through shared memory, so the result is deterministic.
constexpr int COMBINE_GROUPS = 8;
__global__ void __launch_bounds__(128 * COMBINE_GROUPS) pool_combine_kernel(const PoolArgs a)
{
    __shared__ float s_sum[COMBINE_GROUPS][128], s_max[COMBINE_GROUPS][128];
    const int fl = threadIdx.x & 127, grp = threadIdx.x >> 7;
    for (int g = blockIdx.x; g < a.n_graphs; g += gridDim.x) {
        const int64_t n = __ldg(a.node_ptr + g + 1) - __ldg(a.node_ptr + g);
        for (int f0 = 0; f0 < a.F; f0 += 128) {
            const int f = f0 + fl;
            float sum = 0.0f, mx = -INFINITY;
            if (f < a.F) {
                for (int sp = grp; sp < a.splits; sp += COMBINE_GROUPS) {
                    const float *p = a.partial + ((size_t)g * a.splits + sp) * 2 * a.F;
                    sum += p[f];
                    const float v = p[a.F + f];
                    mx = v > mx ? v : mx;
                }
            }
            s_sum[grp][fl] = sum;
            s_max[grp][fl] = mx;
            __syncthreads();
            if (grp == 0 && f < a.F) {
                float ts = 0.0f, tm = -INFINITY;
#pragma unroll
                for (int q = 0; q < COMBINE_GROUPS; q++) {
                    ts += s_sum[q][fl];
                    tm = s_max[q][fl] > tm ? s_max[q][fl] : tm;
                }
                // a NaN in the graph's first row sticks, exactly like the single-pass rule
                const float p0 = a.partial[((size_t)g * a.splits) * 2 * a.F + a.F + f];
                if (p0 != p0) tm = p0;
                write_pools(a, g, f, ts, tm, n);
            }
            __syncthreads();
        }
    }
}

__global__ void activation_kernel(int act, const float *__restrict__ x, float *__restrict__ y,
                                  size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        y[i] = act_apply(act, x[i]);
}

}  // namespace

int launch_pool(const float *x, int ldx, int F, const int64_t *node_ptr, int64_t node_base,
                int n_graphs, int64_t total_nodes, const int *pools, int num_pools, float *pooled,
                DeviceBuf &tmp, cudaStream_t s, int *launches, bool single_pass)
{
    if (n_graphs <= 0 || F <= 0) return GNNB_OK;
    GNNB_REQUIRE(num_pools >= 1 && num_pools <= 4, "pooling: 1..4 aggregations");
    PoolArgs a;
    a.x = x; a.ldx = ldx; a.F = F; a.node_ptr = node_ptr; a.node_base = node_base;
    a.n_graphs = n_graphs;
    a.num_pools = num_pools;
    for (int p = 0; p < 4; p++) a.pools[p] = p < num_pools ? pools[p] : 0;
    a.pooled = pooled;
    // cut graphs that are much larger than the machine into row ranges
    int64_t avg = total_nodes / n_graphs;
    int splits = 1;
    if (avg > 8192 && !single_pass) {   // (STRICT math: one range, the reference's sum order)
        int64_t want = avg / 2048;
        int64_t room = (int64_t)kNumSMs * 16 / n_graphs;
        if (room < 1) room = 1;
        splits = (int)(want < room ? want : room);
        if (splits < 1) splits = 1;
    }
    a.splits = splits;
    a.partial = nullptr;
    if (splits > 1) {
        GNNB_TRY(tmp.ensure(sizeof(float) * (size_t)n_graphs * splits * 2 * F));
        a.partial = tmp.as<float>();
    }
    const int gx = n_graphs < kNumSMs * 32 ? n_graphs : kNumSMs * 32;
    pool_kernel<<<dim3(gx, splits), 128, 0, s>>>(a);
    GNNB_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    if (splits > 1) {
        pool_combine_kernel<<<gx, 128 * COMBINE_GROUPS, 0, s>>>(a);
        GNNB_CUDA(cudaGetLastError());
        if (launches) ++*launches;
    }
    return GNNB_OK;
}

int launch_activation(int act, const float *x, float *y, size_t n, cudaStream_t s, int *launches)
{
    if (n == 0) return GNNB_OK;
    size_t grid = (n + 255) / 256;
    if (grid > (size_t)kNumSMs * 16) grid = (size_t)kNumSMs * 16;
    activation_kernel<<<(int)grid, 256, 0, s>>>(act, x, y, n);
    GNNB_CUDA(cudaGetLastError());
    if (launches) ++*launches;
    return GNNB_OK;
}

}  // namespace gnnb
