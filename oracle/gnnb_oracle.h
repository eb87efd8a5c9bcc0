/*
 * gnnb_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, runtime dimensions) of the float-mode algorithms of the
 * reference header gnnbuilder/gnn_builder_lib/gnn_builder_lib.h ("lib") and of the
 * generated top function gnnbuilder/templates/model.cpp.jinja ("cpp").  It is the
 * checker for the CUDA path: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (gnn_builder_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function against the
 * reference's own golden vectors (gnn_builder_lib_test/tb_data, committed under
 * tests/golden/lib_tb/) and tests/test_oracle_vs_ref.py checks it bit-for-bit against the
 * reference's own templates compiled from /root/reference (oracle/_ref/).
 */
#ifndef GNNB_ORACLE_H
#define GNNB_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* activation ids (same numbering as include/gnnb_b200.h) */
enum {
    ORC_ACT_IDENTITY = 0,
    ORC_ACT_RELU = 1,
    ORC_ACT_GELU_TANH = 2, /* nn.GELU maps to the tanh approximation, cpp:167-168 */
    ORC_ACT_SIGMOID = 3,
    ORC_ACT_TANH = 4,
    ORC_ACT_ELU = 5,
    ORC_ACT_HARDTANH = 6,
    ORC_ACT_LEAKYRELU = 7,
    ORC_ACT_GELU_ERF = 8,
    ORC_ACT_SILU = 9,
    ORC_ACT_SOFTSIGN = 10,
    ORC_ACT_SIN = 11,
    ORC_ACT_COS = 12
};

enum { ORC_CONV_GCN = 0, ORC_CONV_GIN = 1, ORC_CONV_SAGE = 2, ORC_CONV_PNA = 3 };
enum { ORC_POOL_ADD = 0, ORC_POOL_MEAN = 1, ORC_POOL_MAX = 2 };

typedef struct orc_model_desc {
    int conv_type;      /* ORC_CONV_* */
    int num_layers;     /* gnn_num_layers (0 allowed: node_emb_out = features) */
    int in_dim;         /* graph_input_feature_dim */
    int hidden_dim;     /* gnn_hidden_dim */
    int out_dim;        /* gnn_output_dim */
    int skip;           /* gnn_skip_connection */
    int gnn_act;        /* ORC_ACT_* */
    float gin_eps;      /* conv.eps literal, cpp:79 */
    float pna_delta;    /* conv.delta_scaler literal, cpp:113 */
    int num_pools;      /* len(global_pooling.aggrs) */
    int pools[4];       /* ORC_POOL_* in list order */
    int mlp_num_linear; /* hidden_layers + 1 */
    int mlp_hidden;     /* MLP.hidden_dim */
    int mlp_out;        /* MLP.out_dim */
    int mlp_act;        /* MLP.activation */
    int out_act;        /* output_activation (0 = None) */
    int gnn_p_in, gnn_p_hidden, gnn_p_out; /* block factors: only change fp rounding order */
    int mlp_p_in, mlp_p_hidden, mlp_p_out;
} orc_model_desc;

float orc_activation(int act, float x);
void orc_apply_activation(int act, const float *x, float *y, int n);

void orc_linear(const float *x, float *y, const float *W, const float *b, int in_size,
                int out_size, int block_in);

void orc_compute_degree_tables(const int *edge_list, int *in_deg, int *out_deg, int num_nodes,
                               int num_edges);
void orc_compute_neighbor_tables(const int *edge_list, const int *in_deg, int *offsets,
                                 int *neighbor_table, int num_nodes, int num_edges);
void orc_compute_neighbor_and_edge_index_tables(const int *edge_list, const int *in_deg,
                                                int *offsets, int *neighbor_table,
                                                int *edge_index_table, int num_nodes,
                                                int num_edges);

void orc_gcn_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                  const int *in_deg, const float *W, const float *b, int f_in, int f_out,
                  int p_in);
void orc_gin_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                  const int *in_deg, const float *W0, const float *b0, const float *W1,
                  const float *b1, float eps, int f_in, int hidden, int f_out, int p_in);
void orc_gine_conv(int num_nodes, const float *x, float *y, const float *edge_feat,
                   const int *offsets, const int *nbr, const int *edge_index_table,
                   const int *in_deg, const float *We, const float *be, const float *W0,
                   const float *b0, const float *W1, const float *b1, float eps, int f_in,
                   int hidden, int f_out, int f_edge, int p_in);
void orc_sage_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                   const int *in_deg, const float *Wl, const float *bl, const float *Wr, int f_in,
                   int f_out, int p_in);
void orc_pna_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                  const int *in_deg, const float *Wpre, const float *bpre, const float *Wpost,
                  const float *bpost, const float *Wlin, const float *blin, float avg_degree_log,
                  int f_in, int f_out, int p_in, int p_out);
void orc_lg_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                 const int *in_deg, int f);
void orc_simple_conv(int num_nodes, const float *x, float *y, const int *offsets, const int *nbr,
                     const int *in_deg, int f);

void orc_global_add_pool(int num_nodes, const float *x, int f, float *out);
void orc_global_mean_pool(int num_nodes, const float *x, int f, float *out);
void orc_global_max_pool(int num_nodes, const float *x, int f, float *out);

/* number of parameter arrays the model expects, in the reference's flat order
 * (models.py:607-624: mlp_head first, then gnn_convs) */
int orc_model_num_params(const orc_model_desc *d);

/* whole-model forward for one graph, cpp:686-766.  params[] in flat reference order.
 * node_emb_out (nullable) receives [num_nodes][out_dim].  returns 0 on success. */
int orc_model_forward(const orc_model_desc *d, const float *const *params, const float *x,
                      const int *edge_list, int num_nodes, int num_edges, float *out,
                      float *node_emb_out);

/* batch of graphs concatenated: node_ptr[g]..node_ptr[g+1], edge_ptr likewise; edge ids are
 * LOCAL to each graph (as each reference call sees them).  out is [n_graphs][mlp_out]. */
int orc_model_forward_batch(const orc_model_desc *d, const float *const *params, const float *x,
                            const int *edge_list, const long long *node_ptr,
                            const long long *edge_ptr, int n_graphs, float *out);

#ifdef __cplusplus
}
#endif
#endif
