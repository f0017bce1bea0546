/* oracle/gcn_oracle.h — TEST INFRASTRUCTURE (the parity checker), never shipped, never a fallback.
 *
 * Plain-C restatement of the reference's sequential CPU algorithm for the full-batch GCN training
 * path (hengdashi/cuda_gcn, src/seq + src/common).  Every function cites the reference file:line it
 * follows.  The restatement is pinned bit-for-bit against the unmodified reference compiled into
 * oracle/_ref/libgcnref.so (tests/test_oracle_vs_ref.py) and against the committed vectors under
 * tests/golden/ that were produced by that same reference build (tools/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.
 */
#ifndef GCN_ORACLE_H
#define GCN_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG: src/seq/rand.cpp:5-28 ---- */
void     gcno_init_rand_state(long seed);            /* srand(seed); x=rand(); y=rand() until both non-zero */
void     gcno_set_rand_state(uint64_t s0, uint64_t s1);
void     gcno_get_rand_state(uint64_t *out2);
uint32_t gcno_rand(void);                             /* xorshift128plus(), 31-bit */

/* ---- Variable::glorot: src/seq/variable.cpp:11-18 ---- */
void gcno_glorot(float *w, int in_size, int out_size);

/* ---- Modules: src/seq/module.cpp ---- */
void gcno_matmul_fw(const float *a, const float *b, float *c, int m, int n, int p);                 /* :11-22  */
void gcno_matmul_bw(const float *a, const float *b, const float *c_grad, float *a_grad, float *b_grad,
                    int m, int n, int p);                                                          /* :24-42  */
void gcno_spmm_fw(const int *indptr, const int *indices, const float *values, const float *b, float *c,
                  int m, int n, int p);                                                            /* :47-61  */
void gcno_spmm_bw(const int *indptr, const int *indices, const float *values, const float *c_grad,
                  float *b_grad, int m, int n, int p);                                             /* :63-77  */
void gcno_graphsum(const int *indptr, const int *indices, int n, int dim, const float *in, float *out); /* :83-119 (fw and bw share the loop) */
float gcno_cross_entropy(float *logits, const int *truth, float *grad, int n, int num_classes, int training); /* :124-161 */
void gcno_relu_fw(float *x, unsigned char *mask, int n, int training);                              /* :175-185 */
void gcno_relu_bw(float *grad, const unsigned char *mask, int n);                                   /* :187-194 */
void gcno_dropout_fw(float *x, int *mask, int n, float p, int training);                            /* :207-221 (mask may be NULL) */
void gcno_dropout_bw(float *grad, const int *mask, int n, float p);                                 /* :223-233 */

/* ---- Adam: src/seq/optim.cpp:24-37 ---- */
typedef struct gcno_adam gcno_adam;
gcno_adam *gcno_adam_create(int nvars, const int *sizes, const int *decay, float lr, float beta1, float beta2,
                            float eps, float weight_decay);
void gcno_adam_step(gcno_adam *opt, float **data, float **grad);
void gcno_adam_destroy(gcno_adam *opt);

/* ---- GCN helpers: src/seq/gcn.cpp ---- */
void  gcno_set_truth(int *truth, const int *split, const int *label, int n, int current_split);    /* :78-81  */
float gcno_accuracy(const float *logits, const int *truth, int n, int num_classes, int *wrong, int *total); /* :83-96 */
float gcno_l2_penalty(const float *w, int size, float weight_decay);                               /* :98-105 */

/* ---- Parser: src/common/parser.cpp:20-119.  Caller frees the arrays with gcno_data_free. ---- */
typedef struct {
    int num_nodes, input_dim, output_dim;
    long graph_nnz, feature_nnz, n_label, n_split;
    int *graph_indptr, *graph_indices;
    int *feature_indptr, *feature_indices;
    float *feature_value;
    int *label, *split;
} gcno_data;
int  gcno_parse(gcno_data *d, const char *dir, const char *name);   /* 1 ok, 0 if a file cannot be opened */
void gcno_data_free(gcno_data *d);

/* ---- The whole model and loop: src/seq/gcn.cpp:13-158 ---- */
typedef struct gcno_gcn gcno_gcn;
typedef struct {
    int hidden_dim;
    float dropout, learning_rate, weight_decay;
    int epochs, early_stopping;
} gcno_hparams;
gcno_hparams gcno_default_hparams(void);                                                            /* gcn.cpp:9-11 */
/* arrays are borrowed (must outlive the model), exactly like GCNData* in the reference */
gcno_gcn *gcno_gcn_create(const gcno_data *d, gcno_hparams hp, long seed);
void gcno_gcn_destroy(gcno_gcn *g);
void gcno_gcn_train_epoch(gcno_gcn *g, float *loss, float *acc);                                    /* :107-118 */
void gcno_gcn_eval(gcno_gcn *g, int split, float *loss, float *acc);                                /* :120-128 */
long gcno_gcn_var_size(gcno_gcn *g, int idx);           /* idx as gcn.cpp:21-53: 0 input .. 6 logits */
void gcno_gcn_get_var(gcno_gcn *g, int idx, int grad, float *out);
/* run(): prints the reference's lines to stdout (gcn.cpp:139,147,152,157); returns epochs executed */
int  gcno_gcn_run(gcno_gcn *g);

#ifdef __cplusplus
}
#endif
#endif
