// TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" wrappers that instantiate the REFERENCE's own templates
// (gnnbuilder/gnn_builder_lib/gnn_builder_lib.h, included where it lies under
// /root/reference through the reference's own gnn_builder_lib_test/test.h, which supplies the
// float-mode macro set) at the handful of feature dimensions the parity tests and BASELINE
// configs use.  Built by oracle/build_ref.py into oracle/_ref/libgnnb_ref_layers.so; nothing
// from the reference is copied into this repository.
//
// MAX_NODES/MAX_EDGES only size the reference's per-call stack arrays (`int neighbors[MAX_NODES]`,
// lib:1344) and the array-pointer types; the row stride of every table is the feature dim, so
// flat row-major buffers of any node count <= MAX can be passed.
#include <test.h>

#include <cstdlib>

namespace {
constexpr int MAXN = 4096;
constexpr int MAXE = 65536;
}  // namespace

extern "C" {

int ref_max_nodes() { return MAXN; }
int ref_max_edges() { return MAXE; }

void ref_compute_degree_tables(int *edge_list, int *in_deg, int *out_deg, int n, int e)
{
    compute_degree_tables<MAXN, MAXE>((int(*)[2])edge_list, in_deg, out_deg, n, e);
}

void ref_compute_neighbor_tables(int *edge_list, int *in_deg, int *out_deg, int *offsets, int *nbr,
                                 int n, int e)
{
    compute_neighbor_tables<MAXN, MAXE>((int(*)[2])edge_list, in_deg, out_deg, offsets, nbr, n, e);
}

void ref_compute_neighbor_and_edge_index_tables(int *edge_list, int *in_deg, int *out_deg,
                                                int *offsets, int *nbr, int *eidx, int n, int e)
{
    compute_neighbor_and_edge_index_tables<MAXN, MAXE>((int(*)[2])edge_list, in_deg, out_deg,
                                                       offsets, nbr, eidx, n, e);
}

float ref_activation(int act, float x)
{
    switch (act) {
    case 0: return activation_identity<float>(x);
    case 1: return activation_relu<float>(x);
    case 2: return activation_gelu_approx_tanh<float>(x);
    case 3: return activation_sigmoid<float>(x);
    case 4: return activation_tanh<float>(x);
    case 5: return activation_elu<float>(x);
    case 6: return activation_hardtanh<float>(x);
    case 7: return activation_leakyrelu<float>(x);
    case 8: return activation_gelu<float>(x);
    case 9: return activation_silu<float>(x);
    case 10: return activation_softsign<float>(x);
    case 11: return activation_sin<float>(x);
    case 12: return activation_cos<float>(x);
    }
    return 0.0f / 0.0f;
}

#define REF_POOLS(F)                                                                              \
    void ref_global_add_pool_##F(int n, float *x, float *out)                                     \
    {                                                                                             \
        global_add_pool<MAXN, MAXE, F, float>(n, 0, (float(*)[F])x, out);                         \
    }                                                                                             \
    void ref_global_mean_pool_##F(int n, float *x, float *out)                                    \
    {                                                                                             \
        global_mean_pool<MAXN, MAXE, F, float>(n, 0, (float(*)[F])x, out);                        \
    }                                                                                             \
    void ref_global_max_pool_##F(int n, float *x, float *out)                                     \
    {                                                                                             \
        global_max_pool<MAXN, MAXE, F, float>(n, 0, (float(*)[F])x, out);                         \
    }

#define REF_LINEAR(FI, FO)                                                                        \
    void ref_linear_##FI##_##FO(float *x, float *y, float *W, float *b)                           \
    {                                                                                             \
        linear<FI, FO, 1, 1, float>(x, y, (float(*)[FI])W, b);                                    \
    }                                                                                             \
    void ref_linear_buffered_##FI##_##FO(float *x, float *y, float *W, float *b)                  \
    {                                                                                             \
        linear_buffered<FI, FO, 1, 1, float>(x, y, (float(*)[FI])W, b);                           \
    }

#define REF_CONVS(FI, FO)                                                                         \
    void ref_gcn_conv_##FI##_##FO(int n, int e, float *x, float *y, int *coo, int *off, int *nbr, \
                                  int *ind, int *outd, float *W, float *b)                        \
    {                                                                                             \
        gcn_conv<MAXN, MAXE, FI, FO, float>(n, e, (float(*)[FI])x, (float(*)[FO])y,               \
                                            (int(*)[2])coo, off, nbr, ind, outd,                  \
                                            (float(*)[FI])W, b);                                  \
    }                                                                                             \
    void ref_gin_conv_##FI##_##FO(int n, int e, float *x, float *y, int *coo, int *off, int *nbr, \
                                  int *ind, int *outd, float *W0, float *b0, float *W1,           \
                                  float *b1, float eps)                                           \
    {                                                                                             \
        gin_conv<MAXN, MAXE, FI, FO, FO, float>(n, e, (float(*)[FI])x, (float(*)[FO])y,           \
                                                (int(*)[2])coo, off, nbr, ind, outd,              \
                                                (float(*)[FI])W0, b0, (float(*)[FO])W1, b1, eps); \
    }                                                                                             \
    void ref_sage_conv_##FI##_##FO(int n, int e, float *x, float *y, int *coo, int *off,          \
                                   int *nbr, int *ind, int *outd, float *Wl, float *bl,           \
                                   float *Wr)                                                     \
    {                                                                                             \
        /* sage_conv keeps two [MAX_NODES][F] copies on the stack (lib:2239-2240) */              \
        sage_conv<MAXN, MAXE, FI, FO, float>(n, e, (float(*)[FI])x, (float(*)[FO])y,              \
                                             (int(*)[2])coo, off, nbr, ind, outd,                 \
                                             (float(*)[FI])Wl, bl, (float(*)[FI])Wr);             \
    }                                                                                             \
    void ref_pna_conv_##FI##_##FO(int n, int e, float *x, float *y, int *coo, int *off, int *nbr, \
                                  int *ind, int *outd, float *Wpre, float *bpre, float *Wpost,    \
                                  float *bpost, float *Wlin, float *blin, float delta)            \
    {                                                                                             \
        pna_conv<MAXN, MAXE, FI, FO, 2 * FI, FI, 13 * FI, FO, float>(                             \
            n, e, (float(*)[FI])x, (float(*)[FO])y, (int(*)[2])coo, off, nbr, ind, outd,          \
            (float(*)[2 * FI])Wpre, bpre, (float(*)[13 * FI])Wpost, bpost, (float(*)[FO])Wlin,    \
            blin, delta);                                                                         \
    }

#define REF_SAME(F)                                                                               \
    void ref_lg_conv_##F(int n, int e, float *x, float *y, int *coo, int *off, int *nbr,          \
                         int *ind, int *outd)                                                     \
    {                                                                                             \
        lg_conv<MAXN, MAXE, F, F, float>(n, e, (float(*)[F])x, (float(*)[F])y, (int(*)[2])coo,    \
                                         off, nbr, ind, outd);                                    \
    }                                                                                             \
    void ref_simple_conv_##F(int n, int e, float *x, float *y, int *coo, int *off, int *nbr,      \
                             int *ind, int *outd)                                                 \
    {                                                                                             \
        simple_conv<MAXN, MAXE, F, F, float>(n, e, (float(*)[F])x, (float(*)[F])y,                \
                                             (int(*)[2])coo, off, nbr, ind, outd);                \
    }

void ref_gine_conv_8_8_16(int n, int e, float *x, float *y, float *ef, int *coo, int *off,
                          int *nbr, int *eidx, int *ind, int *outd, float *We, float *be,
                          float *W0, float *b0, float *W1, float *b1, float eps)
{
    gine_conv<MAXN, MAXE, 8, 8, 8, 16, float>(n, e, (float(*)[8])x, (float(*)[8])y,
                                              (float(*)[16])ef, (int(*)[2])coo, off, nbr, eidx,
                                              ind, outd, (float(*)[16])We, be, (float(*)[8])W0, b0,
                                              (float(*)[8])W1, b1, eps);
}

REF_CONVS(8, 8)
REF_CONVS(9, 64)
REF_CONVS(64, 64)
REF_CONVS(11, 128)
REF_CONVS(128, 128)
REF_CONVS(9, 128)
REF_CONVS(9, 80)
REF_CONVS(80, 80)
REF_CONVS(5, 12)
REF_CONVS(12, 12)
REF_SAME(8)
REF_POOLS(8)
REF_POOLS(12)
REF_POOLS(64)
REF_POOLS(80)
REF_POOLS(128)
REF_LINEAR(10, 20)
REF_LINEAR(8, 8)
REF_LINEAR(384, 64)
REF_LINEAR(64, 64)
REF_LINEAR(64, 19)
REF_LINEAR(1040, 80)

// ------------------------------------------------------------------ large-graph harness (C5)
// the generated top cannot hold a 2M-node graph in its static arrays, so the
// reference's table builders and gcn_conv are driven directly on heap buffers.  num_rows lets a
// caller time a bounded prefix of destination rows of the same graph.
constexpr int BIGN = 2000000;
constexpr int BIGE = 40000000;

void ref_big_tables(int *edge_list, int *in_deg, int *out_deg, int *offsets, int *nbr, int n, int e)
{
    compute_degree_tables<BIGN, BIGE>((int(*)[2])edge_list, in_deg, out_deg, n, e);
    compute_neighbor_tables<BIGN, BIGE>((int(*)[2])edge_list, in_deg, out_deg, offsets, nbr, n, e);
}

void ref_big_gcn_conv_128_128(int num_rows, int e, float *x, float *y, int *coo, int *off,
                              int *nbr, int *ind, int *outd, float *W, float *b)
{
    gcn_conv<BIGN, BIGE, 128, 128, float>(num_rows, e, (float(*)[128])x, (float(*)[128])y,
                                          (int(*)[2])coo, off, nbr, ind, outd,
                                          (float(*)[128])W, b);
}

}  // extern "C"
