/* include/gcn_host.h — the C face of the C++ host layer (libgcnhost.so): the training loop of
 * hengdashi/cuda_gcn behind plain C, for hosts that cannot link C++ (ctypes, cgo, JNI).
 *
 * The C++ classes in cuda_gcn_b200/host mirror the reference's own API (Variable, SparseIndex, the
 * six Modules, Adam, Parser, GCNParams/GCNData/GCN — reference src/seq/{variable,sparse,module,optim,
 * gcn}.h and src/common/parser.h) and are what a C++ user of the reference switches to.  This header
 * exposes the same objects to non-C++ callers:
 *   gcnh_data     <-> GCNData  (gcn.h:16-22)      gcnh_params <-> GCNParams (gcn.h:9-14)
 *   gcnh_engine   <-> GCN / CUDAGCN (gcn.h:24-44, cuda_gcn.cuh:8-34)
 * Errors follow the reference's convention: a CUDA failure prints "CUDA_ASSERT: ..." and exits
 * (cuda_kernel.cuh:11-18); recoverable conditions (unreadable input) return 0 / NULL.
 * There is no CPU engine: creating an engine without a CUDA device is fatal.
 */
#ifndef GCN_HOST_H
#define GCN_HOST_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int num_nodes, input_dim, hidden_dim, output_dim;
    float dropout, learning_rate, weight_decay;
    int epochs, early_stopping;
} gcnh_params;

typedef struct gcnh_data gcnh_data;
typedef struct gcnh_engine gcnh_engine;

gcnh_params gcnh_default_params(void);                      /* GCNParams::get_default, gcn.cpp:9-11 */

/* ---- GCNData ---- */
gcnh_data *gcnh_data_new(void);
void       gcnh_data_free(gcnh_data *d);
/* Parser(params, data, name).parse() on <root>/<name>.{graph,split,svmlight} (root NULL = "data/");
 * fills params->{num_nodes,input_dim,output_dim}; returns 1 on success, 0 if a file cannot be read
 * or is malformed.  quiet != 0 suppresses the "Parse ... Succeeded." lines. */
int gcnh_data_parse(gcnh_data *d, const char *root, const char *name, gcnh_params *params, int quiet);
/* copy caller arrays in (the reference's GCNData fields are public vectors, gcn.h:16-22) */
int gcnh_data_fill(gcnh_data *d, int num_nodes, const int *graph_indptr, const int *graph_indices,
                   const int *feature_indptr, const int *feature_indices, const float *feature_value,
                   const int *label, const int *split);
/* seeded synthetic dataset: preset in {cora,citeseer,pubmed,reddit,products}, scale in (0,1] shrinks
 * the node and edge counts; returns 0 on an unknown preset or a degree above the int32-safe 46,340 */
int gcnh_data_synth(gcnh_data *d, const char *preset, double scale, uint64_t seed, gcnh_params *params);
/* The row slice rank `rank` of `world` owns under the nnz-balanced partition (gcnk_partition_rows): a new
 * gcnh_data (free it with gcnh_data_free); *row_begin / *row_end receive the range.  Host-only integer work. */
gcnh_data *gcnh_data_slice(const gcnh_data *d, int rank, int world, int *row_begin, int *row_end);
/* sizes[7] = {num_nodes, graph_nnz, feature_nnz, n_label, n_split, max_degree, feature_rows} */
void gcnh_data_sizes(const gcnh_data *d, int64_t *sizes);
/* borrowed pointers into the host vectors (valid until the data is freed / refilled) */
const int   *gcnh_data_graph_indptr(const gcnh_data *d);
const int   *gcnh_data_graph_indices(const gcnh_data *d);
const int   *gcnh_data_feature_indptr(const gcnh_data *d);
const int   *gcnh_data_feature_indices(const gcnh_data *d);
const float *gcnh_data_feature_value(const gcnh_data *d);
const int   *gcnh_data_label(const gcnh_data *d);
const int   *gcnh_data_split(const gcnh_data *d);

/* ---- GCN ---- */
/* plan: 0 auto, 1 modules (the reference's 8-Module chain, unfused kernels), 2 fused.
 * seed: the value the reference would get from time(NULL) (rand.cpp:7); < 0 = $GCN_SEED or time(NULL).
 * device: CUDA device index.  The data must outlive the engine (as GCNData* in the reference). */
gcnh_engine *gcnh_engine_create(const gcnh_params *params, gcnh_data *data, long seed, int plan, int device);
/* Row-partitioned engine (one process or thread per GPU): every rank passes the SAME full data and seed.
 * rank 0 obtains the 128-byte id from gcnh_comm_unique_id and shares it with the other ranks (any transport);
 * all ranks then call gcnh_engine_create_dist, which is collective (NCCL communicator creation). */
int gcnh_comm_unique_id(void *h_id128);
gcnh_engine *gcnh_engine_create_dist(const gcnh_params *params, gcnh_data *data, long seed, int device, int rank, int world,
                                     const void *h_id128);
/* in-place all-reduce of a small host array over the engine's communicator (sum, or max when op_max != 0):
 * barrier + max-over-ranks timing for benchmarks; a no-op for a single-GPU engine */
void gcnh_engine_allreduce_host(gcnh_engine *e, float *h_values, int count, int op_max);
void gcnh_engine_destroy(gcnh_engine *e);
int  gcnh_engine_plan(const gcnh_engine *e);
void gcnh_engine_train_epoch(gcnh_engine *e, float *loss, float *acc);          /* gcn.cpp:107-118 */
void gcnh_engine_eval(gcnh_engine *e, int split, float *loss, float *acc);      /* gcn.cpp:120-128 */
/* train_epoch followed by eval(split) with one host synchronisation — one epoch as GCN::run times it (gcn.cpp:136-138) */
void gcnh_engine_epoch(gcnh_engine *e, int eval_split, float *train_loss, float *train_acc, float *eval_loss, float *eval_acc);
/* integer outputs of the last pass: labelled rows and wrongly classified rows (bit-exact contract) */
void gcnh_engine_last_counts(const gcnh_engine *e, int *count, int *wrong);
int  gcnh_engine_run(gcnh_engine *e, int quiet);            /* GCN::run, gcn.cpp:130-158; returns epochs executed */
/* re-upload the feature values from a host buffer (asynchronous H2D of feature_nnz floats on the
 * engine's stream) — what CUDAGCN::set_input does before every pass (cuda_gcn.cu:81-83) */
void gcnh_engine_set_input_host(gcnh_engine *e, const float *h_values);
/* gcnh_engine_epoch on the CURRENT input while h_next_values (pinned host memory, gcnh_alloc_pinned) is uploaded into a
 * second feature buffer on a copy stream; the next pass of any kind switches to it (device-side wait on the copy).
 * Pipelines the per-epoch re-upload of cuda_gcn.cu:81-83 under the compute.  h_next_values must stay valid and
 * unchanged until the next pass has started.  [opt-in; bench.py --e2e-prefetch] */
void gcnh_engine_epoch_prefetch(gcnh_engine *e, int eval_split, const float *h_next_values, float *train_loss, float *train_acc,
                                float *eval_loss, float *eval_acc);
int64_t gcnh_engine_var_size(const gcnh_engine *e, int idx);
void gcnh_engine_get_var(gcnh_engine *e, int idx, int grad, float *h_out);

/* ---- timers (timer.h:5-26) ---- */
void  gcnh_timer_enable_gpu(int on);                        /* CUDA-event timing of each op */
void  gcnh_timer_enable_mask(unsigned mask);                /* ... of the ops whose slot bit is set only (0 = off) */
int   gcnh_timer_slot(const char *name);                    /* slot index by name ("gather_full", ...), -1 if unknown */
void  gcnh_timer_reset(void);
float gcnh_timer_total(int slot);                           /* seconds */
int   gcnh_timer_calls(int slot);
const char *gcnh_timer_name(int slot);
int   gcnh_timer_count(void);

/* pinned host memory for the e2e path */
float *gcnh_alloc_pinned(int64_t n_floats);
void   gcnh_free_pinned(float *p);

#ifdef __cplusplus
}
#endif
#endif
