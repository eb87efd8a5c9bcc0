/*
 * gnnb_oracle.c -- TEST INFRASTRUCTURE ONLY (see gnnb_oracle.h).
 *
 * Plain-C restatement of the reference's float-mode algorithms.  Every function cites the
 * reference lines it follows: "lib" = gnnbuilder/gnn_builder_lib/gnn_builder_lib.h,
 * "cpp" = gnnbuilder/templates/model.cpp.jinja, "hdr" = gnnbuilder/templates/model.h.jinja.
 *
 * The operation ORDER of every floating-point expression follows the reference so that,
 * built without FMA contraction (-ffp-contract=off, baseline x86-64), results are
 * bit-identical to the reference templates compiled with g++ -O3 (checked by
 * tests/test_oracle_vs_ref.py).  Dimensions are runtime ints instead of template ints and
 * arrays are flat row-major instead of T[MAX][DIM]; nothing else is changed.
 */
#include "gnnb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- activations */

/* lib:308-480 with the float macro set of hdr:18-36 (m_exp=std::exp, m_tanh=std::tanh ...) */
float orc_activation(int act, float x)
{
    switch (act) {
    case ORC_ACT_IDENTITY: /* lib:474-480 */
        return x;
    case ORC_ACT_RELU: /* lib:362-375 */
        return (x > 0) ? x : 0.0f;
    case ORC_ACT_GELU_TANH: { /* lib:387-417 */
        const float GELU_APPROX_MIN = (float)-8.31776613691702;
        const float GELU_TANH_COEFF_LINEAR = (float)0.7978845608028654;
        const float GELU_TANH_COEFF_CUBIC = (float)0.035677408136300125;
        const float GELU_APPROX_MAX = (float)8.31776613691702;
        if (x < GELU_APPROX_MIN)
            return 0.0f;
        if (x > GELU_APPROX_MAX)
            return x;
        const float t0 = GELU_TANH_COEFF_CUBIC * x;
        const float t1 = t0 * x;
        const float t2 = t1 + GELU_TANH_COEFF_LINEAR;
        const float tanh_arg = t2 * x;
        const float th = tanhf(tanh_arg);
        const float th_fixed = (signbit(tanh_arg) != signbit(th)) ? -th : th;
        return (x / 2.0f) * (1.0f + th_fixed);
    }
    case ORC_ACT_SIGMOID: /* lib:419-425 */
        return 1.0f / (1.0f + expf(-x));
    case ORC_ACT_TANH: /* lib:435-448 */
        return tanhf(x);
    case ORC_ACT_ELU: /* lib:308-322 */
        return (x > 0) ? x : 1.0f * (expf(x) - 1.0f);
    case ORC_ACT_HARDTANH: /* lib:324-343 */
        return (x < -1.0f) ? -1.0f : ((x > 1.0f) ? 1.0f : x);
    case ORC_ACT_LEAKYRELU: /* lib:345-360 */
        return (x >= 0) ? x : x * 0.1f;
    case ORC_ACT_GELU_ERF: { /* lib:377-385 */
        const float sqrt_2_recip = 1.0f / sqrtf(2.0f);
        return x * 0.5f * (1.0f + erff(x * sqrt_2_recip));
    }
    case ORC_ACT_SILU: /* lib:427-433 */
        return x * (1.0f / (1.0f + expf(-x)));
    case ORC_ACT_SOFTSIGN: /* lib:450-456 */
        return x / (1.0f + fabsf(x));
    case ORC_ACT_SIN: /* lib:458-464 */
        return sinf(x);
    case ORC_ACT_COS: /* lib:466-472 */
        return cosf(x);
    default:
        return NAN;
    }
}

/* lib:501-509 apply_activation_1d */
void orc_apply_activation(int act, const float *x, float *y, int n)
{
    for (int i = 0; i < n; i++)
        y[i] = orc_activation(act, x[i]);
}

/* ---------------------------------------------------------------- linear */

/* lib:808-905 linear / lib:908-1003 linear_buffered (same arithmetic; the buffered variant
 * only stages the output).  y = bias, then for every block of `block_in` inputs a temp sum
 * starting at 0 is accumulated left to right and added to y.  BLOCK_SIZE_OUT does not change
 * the arithmetic, so it is not a parameter. */
void orc_linear(const float *x, float *y, const float *W, const float *b, int in_size,
                int out_size, int block_in)
{
    if (block_in < 1)
        block_in = 1;
    for (int i = 0; i < out_size; i++)
        y[i] = b ? b[i] : 0.0f; /* lib:852-864 (sage passes a zero bias, lib:2320-2326) */
    for (int i = 0; i < out_size; i++) {
        for (int j = 0; j < in_size; j += block_in) {
            float temp_sum = 0; /* lib:875-880 */
            for (int l = 0; l < block_in && (j + l) < in_size; l++)
                temp_sum += W[(size_t)i * in_size + j + l] * x[j + l]; /* lib:891 */
            y[i] += temp_sum; /* lib:901 */
        }
    }
}

/* ---------------------------------------------------------------- graph tables */

/* lib:1051-1083 */
void orc_compute_degree_tables(const int *edge_list, int *in_deg, int *out_deg, int num_nodes,
                               int num_edges)
{
    for (int i = 0; i < num_nodes; i++) {
        in_deg[i] = 0;
        out_deg[i] = 0;
    }
    for (int i = 0; i < num_edges; i++) {
        int source = edge_list[2 * i + 0];
        int dest = edge_list[2 * i + 1];
        in_deg[dest]++;
        out_deg[source]++;
    }
}

/* lib:1086-1124: offsets = exclusive scan of in-degree (n entries, no sentinel); neighbors
 * placed in COO order => CSR-by-destination, stable. */
void orc_compute_neighbor_tables(const int *edge_list, const int *in_deg, int *offsets,
                                 int *neighbor_table, int num_nodes, int num_edges)
{
    orc_compute_neighbor_and_edge_index_tables(edge_list, in_deg, offsets, neighbor_table, NULL,
                                               num_nodes, num_edges);
}

/* lib:1126-1166 */
void orc_compute_neighbor_and_edge_index_tables(const int *edge_list, const int *in_deg,
                                                int *offsets, int *neighbor_table,
                                                int *edge_index_table, int num_nodes,
                                                int num_edges)
{
    if (num_nodes <= 0)
        return;
    int *cursor = (int *)malloc(sizeof(int) * (size_t)num_nodes);
    offsets[0] = 0;
    cursor[0] = 0;
    for (int i = 1; i < num_nodes; i++) {
        int csum = offsets[i - 1] + in_deg[i - 1];
        offsets[i] = csum;
        cursor[i] = csum;
    }
    for (int i = 0; i < num_edges; i++) {
        int source = edge_list[2 * i + 0];
        int dest = edge_list[2 * i + 1];
        int offset = cursor[dest];
        neighbor_table[offset] = source;
        if (edge_index_table)
            edge_index_table[offset] = i;
        cursor[dest]++;
    }
    free(cursor);
}

/* ---------------------------------------------------------------- conv layers */

/* lib:1213-1289 gcn_conv_agg + lib:1291-1387 gcn_conv */
void orc_gcn_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                  const int *in_deg, const float *W, const float *b, int f_in, int f_out,
                  int p_in)
{
    float *agg = (float *)malloc(sizeof(float) * (size_t)f_in);
    for (int node = 0; node < num_nodes; node++) {
        const int deg = in_deg[node];
        const int off = offsets[node];
        for (int i = 0; i < f_in; i++)
            agg[i] = 0.0f;
        for (int k = 0; k < deg; k++) {
            const int u = nbr[off + k];
            const float d_i_prime = 1.0f + (float)deg;
            const float d_j_prime = 1.0f + (float)in_deg[u];
            const float scale = 1.0f / sqrtf(d_i_prime * d_j_prime); /* lib:1252 */
            for (int i = 0; i < f_in; i++)
                agg[i] += x[(size_t)u * f_in + i] * scale; /* lib:1257,1262 */
        }
        const float d_i_prime = 1.0f + deg;                             /* lib:1266 */
        const float scale_self = 1.0f / sqrtf(d_i_prime * d_i_prime);   /* lib:1267 */
        for (int i = 0; i < f_in; i++)
            agg[i] += x[(size_t)node * f_in + i] * scale_self; /* lib:1272,1277 */
        orc_linear(agg, y + (size_t)node * f_out, W, b, f_in, f_out, p_in); /* lib:1379 */
    }
    free(agg);
}

/* shared tail of gin/gine: lib:1519-1547 */
static void gin_apply(const float *agg, const float *self, float *out, const float *W0,
                      const float *b0, const float *W1, const float *b1, float eps, int f_in,
                      int hidden, int f_out, int p_in, float *h_in, float *h_mid)
{
    for (int i = 0; i < f_in; i++) {
        const float self_scaled = self[i] * (1.0f + eps); /* lib:1522 */
        h_in[i] = agg[i] + self_scaled;                   /* lib:1528 */
    }
    orc_linear(h_in, h_mid, W0, b0, f_in, hidden, p_in); /* lib:1537 */
    for (int i = 0; i < hidden; i++)
        h_mid[i] = orc_activation(ORC_ACT_RELU, h_mid[i]); /* lib:1540 */
    orc_linear(h_mid, out, W1, b1, hidden, f_out, p_in);   /* lib:1542 */
}

/* lib:1389-1437 gin_conv_agg + lib:1440-1549 gin_conv */
void orc_gin_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                  const int *in_deg, const float *W0, const float *b0, const float *W1,
                  const float *b1, float eps, int f_in, int hidden, int f_out, int p_in)
{
    float *agg = (float *)malloc(sizeof(float) * (size_t)(2 * f_in + hidden));
    float *h_in = agg + f_in, *h_mid = h_in + f_in;
    for (int node = 0; node < num_nodes; node++) {
        const int deg = in_deg[node];
        const int off = offsets[node];
        for (int i = 0; i < f_in; i++)
            agg[i] = 0.0f;
        for (int k = 0; k < deg; k++) {
            const int u = nbr[off + k];
            for (int i = 0; i < f_in; i++)
                agg[i] += x[(size_t)u * f_in + i]; /* lib:1424 */
        }
        gin_apply(agg, x + (size_t)node * f_in, y + (size_t)node * f_out, W0, b0, W1, b1, eps,
                  f_in, hidden, f_out, p_in, h_in, h_mid);
    }
    free(agg);
}

/* lib:1555-1623 gine_conv_agg + lib:1627-1742 gine_conv */
void orc_gine_conv(int num_nodes, const float *x, float *y, const float *edge_feat,
                   const int *offsets, const int *nbr, const int *edge_index_table,
                   const int *in_deg, const float *We, const float *be, const float *W0,
                   const float *b0, const float *W1, const float *b1, float eps, int f_in,
                   int hidden, int f_out, int f_edge, int p_in)
{
    float *agg = (float *)malloc(sizeof(float) * (size_t)(3 * f_in + hidden));
    float *h_in = agg + f_in, *proj = h_in + f_in, *h_mid = proj + f_in;
    for (int node = 0; node < num_nodes; node++) {
        const int deg = in_deg[node];
        const int off = offsets[node];
        for (int i = 0; i < f_in; i++)
            agg[i] = 0.0f;
        for (int k = 0; k < deg; k++) {
            const int u = nbr[off + k];
            const int eid = edge_index_table[off + k];
            orc_linear(edge_feat + (size_t)eid * f_edge, proj, We, be, f_edge, f_in,
                       p_in); /* lib:1600 */
            for (int i = 0; i < f_in; i++) {
                const float message = x[(size_t)u * f_in + i] + proj[i];   /* lib:1603 */
                agg[i] += orc_activation(ORC_ACT_RELU, message);            /* lib:1606,1610 */
            }
        }
        gin_apply(agg, x + (size_t)node * f_in, y + (size_t)node * f_out, W0, b0, W1, b1, eps,
                  f_in, hidden, f_out, p_in, h_in, h_mid);
    }
    free(agg);
}

/* lib:2161-2209 sage_conv_agg (mean_incremental, lib:646-669) + lib:2211-2341 sage_conv */
void orc_sage_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                   const int *in_deg, const float *Wl, const float *bl, const float *Wr, int f_in,
                   int f_out, int p_in)
{
    float *sum = (float *)malloc(sizeof(float) * (size_t)(2 * f_in + 2 * f_out));
    float *mean = sum + f_in, *a = mean + f_in, *s = a + f_out;
    for (int node = 0; node < num_nodes; node++) {
        const int deg = in_deg[node];
        const int off = offsets[node];
        for (int i = 0; i < f_in; i++) {
            sum[i] = 0.0f;
            mean[i] = 0.0f; /* stays 0 with no samples, lib:651 */
        }
        int count = 0;
        for (int k = 0; k < deg; k++) {
            const int u = nbr[off + k];
            count++;
            for (int i = 0; i < f_in; i++) {
                sum[i] += x[(size_t)u * f_in + i]; /* lib:659 */
                mean[i] = sum[i] / (float)count;   /* lib:661 */
            }
        }
        orc_linear(mean, a, Wl, bl, f_in, f_out, p_in);                         /* lib:2316 */
        orc_linear(x + (size_t)node * f_in, s, Wr, NULL, f_in, f_out, p_in);    /* lib:2326 */
        for (int i = 0; i < f_out; i++)
            y[(size_t)node * f_out + i] = a[i] + s[i]; /* lib:2332 */
    }
    free(sum);
}

/* lib:1750-1834 pna_conv_agg, lib:1836-1876 pna_conv_concat, lib:1891-2157 pna_conv.
 * accumulators: max/min lib:736-802, mean lib:646-669, Welford variance lib:677-705. */
void orc_pna_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                  const int *in_deg, const float *Wpre, const float *bpre, const float *Wpost,
                  const float *bpost, const float *Wlin, const float *blin, float avg_degree_log,
                  int f_in, int f_out, int p_in, int p_out)
{
    const int F = f_in;
    const int concat = 13 * F;
    float *buf = (float *)malloc(sizeof(float) * (size_t)(2 * F + F + 8 * F + concat + f_out));
    float *cat2 = buf;            /* [2F] self || neighbor */
    float *t = cat2 + 2 * F;      /* [F] transformed */
    float *vmax = t + F, *vmin = vmax + F, *vsum = vmin + F, *vmean = vsum + F;
    float *wmean = vmean + F, *wm2 = wmean + F, *vstd = wm2 + F, *spare = vstd + F;
    float *pre = spare + F;       /* [13F] */
    float *hid = pre + concat;    /* [f_out] */
    (void)spare;
    for (int node = 0; node < num_nodes; node++) {
        const int deg = in_deg[node];
        const int off = offsets[node];
        const float *self = x + (size_t)node * F;
        const int clamped = (deg < 1) ? 1 : deg;                                  /* lib:1973-1981 */
        const float amplification = logf((float)(clamped + 1)) / avg_degree_log;  /* lib:1983 */
        const float attenuation = avg_degree_log / logf((float)(clamped + 1));    /* lib:1984 */

        for (int i = 0; i < F; i++) {
            vmax[i] = 0.0f; vmin[i] = 0.0f; vsum[i] = 0.0f; vmean[i] = 0.0f;
            wmean[i] = 0.0f; wm2[i] = 0.0f;
        }
        int count = 0;
        for (int k = 0; k < deg; k++) {
            const int u = nbr[off + k];
            for (int i = 0; i < F; i++) {
                cat2[i] = self[i];                       /* lib:1801 */
                cat2[i + F] = x[(size_t)u * F + i];      /* lib:1802 */
            }
            orc_linear(cat2, t, Wpre, bpre, 2 * F, F, p_in); /* lib:1807 */
            count++;
            for (int i = 0; i < F; i++) {
                const float v = t[i];
                if (count == 1) { /* first sample, lib:748-752 / 784-788 */
                    vmax[i] = v;
                    vmin[i] = v;
                } else {
                    if (v > vmax[i]) vmax[i] = v;
                    if (v < vmin[i]) vmin[i] = v;
                }
                vsum[i] += v;                       /* lib:659 */
                vmean[i] = vsum[i] / (float)count;  /* lib:661 */
                const float delta = v - wmean[i];   /* lib:693 */
                wmean[i] += delta / (float)count;   /* lib:694 */
                wm2[i] += delta * (v - wmean[i]);   /* lib:695 */
            }
        }
        for (int i = 0; i < F; i++) {
            const float var = wm2[i] / (float)count; /* lib:702 (0/0 = NaN when deg == 0) */
            vstd[i] = sqrtf(var + 1e-5f);            /* lib:703 */
        }
        for (int i = 0; i < F; i++) { /* lib:1857-1875 with lib:2081-2089 */
            pre[i + F * 0] = self[i];
            pre[i + F * 1] = vmax[i];
            pre[i + F * 2] = vmin[i];
            pre[i + F * 3] = vmean[i];
            pre[i + F * 4] = vstd[i];
            pre[i + F * 5] = amplification * vmax[i];
            pre[i + F * 6] = amplification * vmin[i];
            pre[i + F * 7] = amplification * vmean[i];
            pre[i + F * 8] = amplification * vstd[i];
            pre[i + F * 9] = attenuation * vmax[i];
            pre[i + F * 10] = attenuation * vmin[i];
            pre[i + F * 11] = attenuation * vmean[i];
            pre[i + F * 12] = attenuation * vstd[i];
        }
        orc_linear(pre, hid, Wpost, bpost, concat, f_out, p_in);                    /* lib:2149 */
        orc_linear(hid, y + (size_t)node * f_out, Wlin, blin, f_out, f_out, p_out); /* lib:2150 */
    }
    free(buf);
}

/* lib:2350-2409 lg_conv_agg + lib:2411-2499 lg_conv (no self term, no linear) */
void orc_lg_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                 const int *in_deg, int f)
{
    for (int node = 0; node < num_nodes; node++) {
        const int deg = in_deg[node];
        const int off = offsets[node];
        float *agg = y + (size_t)node * f;
        for (int i = 0; i < f; i++)
            agg[i] = 0.0f;
        for (int k = 0; k < deg; k++) {
            const int u = nbr[off + k];
            const float scale = 1.0f / sqrtf((float)(deg * in_deg[u])); /* lib:2386 */
            for (int i = 0; i < f; i++)
                agg[i] += x[(size_t)u * f + i] * scale;
        }
    }
}

/* lib:2511-2549 simple_conv_agg + lib:2564-2634 simple_conv */
void orc_simple_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                     const int *in_deg, int f)
{
    for (int node = 0; node < num_nodes; node++) {
        const int deg = in_deg[node];
        const int off = offsets[node];
        float *agg = y + (size_t)node * f;
        for (int i = 0; i < f; i++)
            agg[i] = 0.0f;
        for (int k = 0; k < deg; k++) {
            const int u = nbr[off + k];
            for (int i = 0; i < f; i++)
                agg[i] += x[(size_t)u * f + i];
        }
    }
}

/* ---------------------------------------------------------------- global pooling */

/* lib:2709-2739 */
void orc_global_add_pool(int num_nodes, const float *x, int f, float *out)
{
    for (int j = 0; j < f; j++)
        out[j] = 0.0f;
    for (int i = 0; i < num_nodes; i++)
        for (int j = 0; j < f; j++)
            out[j] += x[(size_t)i * f + j];
}

/* lib:2741-2771 (running mean: the value left after the last update is sum / n) */
void orc_global_mean_pool(int num_nodes, const float *x, int f, float *out)
{
    float *sum = (float *)malloc(sizeof(float) * (size_t)(f > 0 ? f : 1));
    for (int j = 0; j < f; j++) {
        sum[j] = 0.0f;
        out[j] = 0.0f;
    }
    for (int i = 0; i < num_nodes; i++)
        for (int j = 0; j < f; j++) {
            sum[j] += x[(size_t)i * f + j];
            out[j] = sum[j] / (float)(i + 1);
        }
    free(sum);
}

/* lib:2773-2803 (first sample initialises; 0 with no samples) */
void orc_global_max_pool(int num_nodes, const float *x, int f, float *out)
{
    for (int j = 0; j < f; j++)
        out[j] = 0.0f;
    for (int i = 0; i < num_nodes; i++)
        for (int j = 0; j < f; j++) {
            const float v = x[(size_t)i * f + j];
            if (i == 0 || v > out[j])
                out[j] = v;
        }
}

/* ---------------------------------------------------------------- whole model */

static int conv_params_per_layer(int conv_type)
{
    switch (conv_type) {
    case ORC_CONV_GCN: return 2;  /* conv_bias, conv_lin_weight */
    case ORC_CONV_GIN: return 4;  /* mlp_linear_0_{weight,bias}, mlp_linear_1_{weight,bias} */
    case ORC_CONV_SAGE: return 3; /* conv_lin_l_{weight,bias}, conv_lin_r_weight */
    case ORC_CONV_PNA: return 6;  /* pre_nns_0_0_{w,b}, post_nns_0_0_{w,b}, lin_{w,b} */
    default: return -1;
    }
}

int orc_model_num_params(const orc_model_desc *d)
{
    int c = conv_params_per_layer(d->conv_type);
    if (c < 0)
        return -1;
    return 2 * d->mlp_num_linear + c * d->num_layers;
}

static void layer_dims(const orc_model_desc *d, int layer, int *fi, int *fo)
{
    /* models.py:519-549 */
    if (d->num_layers == 1) {
        *fi = d->in_dim;
        *fo = d->out_dim;
        return;
    }
    *fi = (layer == 0) ? d->in_dim : d->hidden_dim;
    *fo = (layer == d->num_layers - 1) ? d->out_dim : d->hidden_dim;
}

static void layer_p(const orc_model_desc *d, int layer, int *pi, int *po)
{
    /* models.py:519-549 p_in/p_out threading */
    if (d->num_layers == 1) {
        *pi = d->gnn_p_in;
        *po = d->gnn_p_out;
        return;
    }
    *pi = (layer == 0) ? d->gnn_p_in : d->gnn_p_hidden;
    *po = (layer == d->num_layers - 1) ? d->gnn_p_out : d->gnn_p_hidden;
}

static int imax(int a, int b) { return a > b ? a : b; }

/* cpp:686-766: tables -> compute_gnn_head (cpp:151-359) -> compute_global_graph_pooling
 * (cpp:413-449) -> compute_mlp_head (cpp:454-530) -> compute_model_output (cpp:627-651) */
int orc_model_forward(const orc_model_desc *d, const float *const *params, const float *x,
                      const int *edge_list, int num_nodes, int num_edges, float *out,
                      float *node_emb_out)
{
    const int cpl = conv_params_per_layer(d->conv_type);
    if (cpl < 0 || num_nodes < 0 || num_edges < 0)
        return 1;
    const int n = num_nodes, e = num_edges;
    int maxf = imax(imax(d->in_dim, d->hidden_dim), d->out_dim);

    int *in_deg = (int *)calloc((size_t)imax(n, 1), sizeof(int));
    int *out_deg = (int *)calloc((size_t)imax(n, 1), sizeof(int));
    int *offsets = (int *)calloc((size_t)imax(n, 1), sizeof(int));
    int *nbr = (int *)calloc((size_t)imax(e, 1), sizeof(int));
    float *cur = (float *)calloc((size_t)imax(n, 1) * maxf, sizeof(float));
    float *nxt = (float *)calloc((size_t)imax(n, 1) * maxf, sizeof(float));

    orc_compute_degree_tables(edge_list, in_deg, out_deg, n, e);          /* cpp:737-743 */
    orc_compute_neighbor_tables(edge_list, in_deg, offsets, nbr, n, e);   /* cpp:745-758 */

    memcpy(cur, x, sizeof(float) * (size_t)n * d->in_dim); /* cpp:257-262 */
    int cur_dim = d->in_dim;
    const float *const *cp = params + 2 * d->mlp_num_linear; /* conv params follow the head's */
    for (int layer = 0; layer < d->num_layers; layer++) {
        int fi, fo, pi, po;
        layer_dims(d, layer, &fi, &fo);
        layer_p(d, layer, &pi, &po);
        const float *const *p = cp + (size_t)layer * cpl;
        switch (d->conv_type) {
        case ORC_CONV_GCN: /* cpp:27-50: weight = conv_lin_weight, bias = conv_bias */
            orc_gcn_conv(n, cur, nxt, offsets, nbr, in_deg, p[1], p[0], fi, fo, pi);
            break;
        case ORC_CONV_GIN: /* cpp:53-80, hidden = out_channels (models.py:52-53,90) */
            orc_gin_conv(n, cur, nxt, offsets, nbr, in_deg, p[0], p[1], p[2], p[3], d->gin_eps, fi,
                         fo, fo, pi);
            break;
        case ORC_CONV_SAGE: /* cpp:116-140 */
            orc_sage_conv(n, cur, nxt, offsets, nbr, in_deg, p[0], p[1], p[2], fi, fo, pi);
            break;
        case ORC_CONV_PNA: /* cpp:82-114 */
            orc_pna_conv(n, cur, nxt, offsets, nbr, in_deg, p[0], p[1], p[2], p[3], p[4], p[5],
                         d->pna_delta, fi, fo, pi, po);
            break;
        }
        const int do_skip = d->skip && layer != 0 && layer != d->num_layers - 1; /* cpp:269-279 */
        for (size_t i = 0; i < (size_t)n * fo; i++) {
            float v = nxt[i];
            if (do_skip)
                v = cur[i] + v; /* cpp:308: in_skip + post_conv (needs fi == fo) */
            nxt[i] = orc_activation(d->gnn_act, v); /* cpp:313-322 */
        }
        float *tmp = cur;
        cur = nxt;
        nxt = tmp;
        cur_dim = fo;
    }
    if (node_emb_out)
        memcpy(node_emb_out, cur, sizeof(float) * (size_t)n * cur_dim);

    /* pooling, cpp:440-448: pools concatenated in list order */
    const int emb = (d->num_layers == 0) ? d->in_dim : d->out_dim;
    const int head_in = emb * d->num_pools;
    int maxh = imax(imax(head_in, d->mlp_hidden), d->mlp_out);
    float *h0 = (float *)calloc((size_t)imax(maxh, 1), sizeof(float));
    float *h1 = (float *)calloc((size_t)imax(maxh, 1), sizeof(float));
    for (int k = 0; k < d->num_pools; k++) {
        float *dst = h0 + (size_t)k * emb;
        switch (d->pools[k]) {
        case ORC_POOL_ADD: orc_global_add_pool(n, cur, emb, dst); break;
        case ORC_POOL_MEAN: orc_global_mean_pool(n, cur, emb, dst); break;
        case ORC_POOL_MAX: orc_global_max_pool(n, cur, emb, dst); break;
        default: return 2;
        }
    }
    /* MLP head, cpp:483-529 + models.py:398-446 */
    int in_f = head_in;
    for (int l = 0; l < d->mlp_num_linear; l++) {
        const int last = (l == d->mlp_num_linear - 1);
        const int out_f = last ? d->mlp_out : d->mlp_hidden;
        int bi;
        if (d->mlp_num_linear == 1)
            bi = d->mlp_p_in;
        else
            bi = (l == 0) ? d->mlp_p_in : d->mlp_p_hidden;
        orc_linear(h0, h1, params[2 * l], params[2 * l + 1], in_f, out_f, bi); /* cpp:496-507 */
        if (!last)
            orc_apply_activation(d->mlp_act, h1, h1, out_f); /* cpp:510-513 */
        float *tmp = h0;
        h0 = h1;
        h1 = tmp;
        in_f = out_f;
    }
    for (int i = 0; i < d->mlp_out; i++) /* cpp:644-650 */
        out[i] = d->out_act ? orc_activation(d->out_act, h0[i]) : h0[i];

    free(in_deg); free(out_deg); free(offsets); free(nbr); free(cur); free(nxt);
    free(h0); free(h1);
    return 0;
}

int orc_model_forward_batch(const orc_model_desc *d, const float *const *params, const float *x,
                            const int *edge_list, const long long *node_ptr,
                            const long long *edge_ptr, int n_graphs, float *out)
{
    for (int g = 0; g < n_graphs; g++) {
        const int n = (int)(node_ptr[g + 1] - node_ptr[g]);
        const int e = (int)(edge_ptr[g + 1] - edge_ptr[g]);
        int rc = orc_model_forward(d, params, x + (size_t)node_ptr[g] * d->in_dim,
                                   edge_list + 2 * (size_t)edge_ptr[g], n, e,
                                   out + (size_t)g * d->mlp_out, NULL);
        if (rc)
            return rc;
    }
    return 0;
}
