// gcn.h — GCNParams / GCNData / GCN with the reference's public shape (src/seq/gcn.h:9-44; GPU twin
// CUDAGCN, src/cuda/cuda_gcn.cuh:30-33): the 2-layer Kipf-Welling GCN training loop that
// `./gcn-cuda <dataset>` runs.  GCN(params, &data).run() prints the reference's lines.
//
// Two execution plans over the same Variables (all resident in HBM, uploaded once):
//   PLAN_MODULES  the reference's chain of 8 Module objects, forward then backward in reverse
//                 (gcn.cpp:107-128), one unfused kernel per operator.  Every intermediate Variable
//                 (V0..V6 of gcn.cpp:21-53) exists and is inspectable: this is the verification plan.
//   PLAN_FUSED    the static fused plan (default when the adjacency is symmetric): dropout-on-read
//                 feature transform, GraphSum with the ReLU/Dropout epilogue, layer 2 re-ordered to
//                 (A_hat*H1)*W2 so that every gather runs at the hidden width, Matmul + softmax-CE +
//                 accuracy + Matmul backward in one row-local kernel, fused backward gathers, one
//                 multi-tensor Adam launch.  Same RNG stream, same printed numbers within fp32 rounding.
#pragma once
#include <memory>
#include <utility>
#include <vector>

#include "module.h"
#include "optim.h"
#include "sparse.h"
#include "variable.h"

struct GCNParams {
    int num_nodes, input_dim, hidden_dim, output_dim;
    float dropout, learning_rate, weight_decay;
    int epochs, early_stopping;
    static GCNParams get_default();
};

class GCNData {
public:
    SparseIndex feature_index, graph;
    std::vector<int> split;
    std::vector<int> label;
    std::vector<float> feature_value;
};

enum GCNPlan { PLAN_AUTO = 0, PLAN_MODULES = 1, PLAN_FUSED = 2 };

// Rows [r0, r0 + n_rows) of a dataset: graph rows with GLOBAL column ids, feature rows, labels, split.
// Pure host integer work; concatenating the slices of a partition reproduces the input bit for bit.
void slice_rows(const GCNData &src, int r0, int n_rows, GCNData &dst);

// Row-partitioned execution (the reference is single-GPU): rank k of `world` owns the contiguous,
// nnz-balanced row range gcnk_partition_rows assigns it.  Every rank is constructed from the SAME full GCNData
// and the same seed; `comm` is an initialised gcnk_comm (include/gcnk.h).  world == 1 is the plain engine.
struct GCNDist {
    int rank = 0, world = 1;
    gcnk_comm *comm = nullptr;
};

class GCN {
public:
    GCN(GCNParams params, GCNData *data);                       // plan from $GCN_PLAN (modules|fused), default auto
    GCN(GCNParams params, GCNData *data, GCNPlan plan, bool quiet);
    GCN(GCNParams params, GCNData *data, GCNPlan plan, bool quiet, GCNDist dist);
    ~GCN();
    GCNParams params;
    void run();

    // ---- beyond the reference's public surface (private there); used by the C face, tests and bench
    std::pair<float, float> train_epoch();
    std::pair<float, float> eval(int current_split);
    // train_epoch() + eval(split) with one host synchronisation (fused plan); the same numbers as the two calls
    void epoch(int eval_split, float *train_loss, float *train_acc, float *eval_loss, float *eval_acc);
    int train_count = 0, train_wrong = 0;                        // integer outputs of the training half of epoch()
    GCNPlan plan() const { return plan_; }
    void set_input_from_host(const float *h_values);            // re-upload the feature values (H2D of nnz(X) floats)
    // epoch() on the current input while `h_next` (pinned host memory) is uploaded into a second feature buffer on a
    // copy stream; the next pass of any kind switches to it.  Pipelines a per-epoch re-upload under the compute.
    void epoch_prefetch(int eval_split, const float *h_next, float *train_loss, float *train_acc, float *eval_loss, float *eval_acc);
    // Variable idx as constructed in gcn.cpp:21-53 (0 input, 1 X*W1, 2 W1, 3 layer-1 out, 4 H1*W2, 5 W2, 6 logits).
    // In the fused plan 4 does not exist (size 0), 1/3/6 are materialised on demand for data only.
    long var_size(int idx) const;
    void get_var(int idx, bool grad, float *h_out);
    int last_count = 0, last_wrong = 0;                          // labelled rows / wrongly classified, last pass
    int epochs_run = 0;

private:
    void build(GCNPlan plan);
    void set_input();
    void set_truth(int current_split);
    float get_accuracy();
    float get_l2_penalty();
    std::pair<float, float> fused_pass(int current_split, bool training);
    void fused_enqueue(int current_split, bool training, int slot);
    std::pair<float, float> fused_collect(int slot, bool sync);

    void build_partition();
    gcnk_graph *graph_handle();
    void build_halo();
    void build_wide();
    void build_wide_halo();
    void finish_build();
    void wide_enqueue(int current_split, bool training, int slot);
    void enqueue_loss_sum(int split_index, bool training, int slot);
    void launch_loss_sum(int sidx_l, bool training);
    void flush_loss_sum();
    void finish_pass(bool training, bool seq, int slot);
    gcnk_stream_t engine_stream() const;
    void mirror(float *d_all, int dim);
    void publish(float *d_all, int dim, bool on_comm_stream = false);
    bool exchange_overlapped(float *buf, int dim, gcnk_graph *v_own, gcnk_graph *v_rem);
    bool arm_exchange(float *d_all, int dim);
    void await(float *d_all, int dim);
    GCNData *data;                           // what this rank computes on: the caller's data, or `local` (its row slice)
    GCNData *full_data = nullptr;            // the caller's full data
    std::unique_ptr<GCNData> local;
    GCNDist dist;
    std::vector<int> row_begin;              // [world + 1] partition cuts
    int n_loc = 0, r0 = 0;                   // rows owned by this rank, first owned row
    float *d_dinv_global = nullptr;          // [N] d^-1/2 of every node (partitioned runs)
    GCNPlan plan_ = PLAN_MODULES;
    bool quiet_ = false;
    std::vector<Module *> modules;
    std::vector<Variable> variables;
    Variable *input = nullptr, *output = nullptr;
    CrossEntropyLoss *ce_module = nullptr;
    int *d_truth = nullptr, *d_split = nullptr, *d_label = nullptr;
    float *d_feature_value = nullptr;        // pristine feature values (never modified)
    float *d_feature_spare = nullptr;        // epoch_prefetch: the buffer the next input is uploaded into
    gcnk_stream_t copy_stream = nullptr;
    void *ev_copied = nullptr;
    bool input_pending = false;              // d_feature_spare holds a newer input (complete when ev_copied has fired)
    void start_input_upload(const float *h_values);
    void consume_pending_input();
    Adam optimizer;
    float loss = 0;
    int split_count[4] = {0, 0, 0, 0};

    // fused-plan state
    struct Fused;
    std::unique_ptr<Fused> fz;
};
