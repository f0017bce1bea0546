#include "module.h"

#include "check.h"
#include "rand.h"
#include "timer.h"

// ---------------------------------------------------------------------------------- Matmul ----
Matmul::Matmul(Variable *a, Variable *b, Variable *c, int m, int n, int p) : a(a), b(b), c(c), m(m), n(n), p(p) {
    workspace_bytes = gcnk_matmul_bw_b_workspace(m, n, p);
    if (workspace_bytes) GCNK_CHECK(gcnk_malloc((void **)&workspace, workspace_bytes));
}
Matmul::~Matmul() { if (workspace) gcnk_free(workspace); }

void Matmul::forward(bool) {
    gpu_timer_begin(TMR_MATMUL_FW);
    GCNK_CHECK(gcnk_matmul_fw(a->data, b->data, c->data, m, n, p, nullptr));
    gpu_timer_end(TMR_MATMUL_FW);
}

void Matmul::backward() {
    gpu_timer_begin(TMR_MATMUL_BW);
    GCNK_CHECK(gcnk_matmul_bw_a(c->grad, b->data, a->grad, m, n, p, nullptr));
    GCNK_CHECK(gcnk_matmul_bw_b(a->data, c->grad, b->grad, m, n, p, workspace, workspace_bytes, nullptr));
    gpu_timer_end(TMR_MATMUL_BW);
}

// ---------------------------------------------------------------------------- SparseMatmul ----
SparseMatmul::SparseMatmul(Variable *a, Variable *b, Variable *c, SparseIndex *sp, int m, int n, int p)
    : a(a), b(b), c(c), sp(sp), m(m), n(n), p(p) {}

void SparseMatmul::forward(bool) {
    gpu_timer_begin(TMR_SPMATMUL_FW);
    GCNK_CHECK(gcnk_spmm_fw(sp->spmat(m, n), a->data, b->data, c->data, p, nullptr, 1.0f, nullptr, nullptr));
    gpu_timer_end(TMR_SPMATMUL_FW);
}

void SparseMatmul::backward() {
    gpu_timer_begin(TMR_SPMATMUL_BW);
    GCNK_CHECK(gcnk_spmm_bw(sp->spmat(m, n), a->data, c->grad, b->grad, p, nullptr, 1.0f, nullptr));
    gpu_timer_end(TMR_SPMATMUL_BW);
}

// -------------------------------------------------------------------------------- GraphSum ----
GraphSum::GraphSum(Variable *in, Variable *out, SparseIndex *graph, int dim) : in(in), out(out), graph(graph), dim(dim) {}

void GraphSum::forward(bool) {
    gpu_timer_begin(TMR_GRAPHSUM_FW);
    GCNK_CHECK(gcnk_graphsum(graph->graph(), in->data, out->data, dim, nullptr));
    gpu_timer_end(TMR_GRAPHSUM_FW);
}

void GraphSum::backward() {
    gpu_timer_begin(TMR_GRAPHSUM_BW);
    GCNK_CHECK(gcnk_graphsum(graph->graph(), out->grad, in->grad, dim, nullptr));
    gpu_timer_end(TMR_GRAPHSUM_BW);
}

// ------------------------------------------------------------------------ CrossEntropyLoss ----
CrossEntropyLoss::CrossEntropyLoss(Variable *logits, int *truth, float *loss, int num_classes)
    : logits(logits), truth(truth), loss(loss), num_classes(num_classes) {
    const int n = logits->size / num_classes;
    workspace_bytes = gcnk_softmax_ce_workspace(n, num_classes);
    GCNK_CHECK(gcnk_malloc((void **)&workspace, workspace_bytes));
    GCNK_CHECK(gcnk_malloc((void **)&d_result, sizeof(gcnk_ce_result)));
}
CrossEntropyLoss::~CrossEntropyLoss() { gcnk_free(workspace); gcnk_free(d_result); }

void CrossEntropyLoss::forward(bool training) {
    const int n = logits->size / num_classes;
    gpu_timer_begin(TMR_LOSS_FW);
    GCNK_CHECK(gcnk_softmax_ce(logits->data, truth, training ? logits->grad : nullptr, n, num_classes, training, d_result,
                               workspace, workspace_bytes, nullptr));
    gpu_timer_end(TMR_LOSS_FW);
    gcnk_ce_result r;
    GCNK_CHECK(gcnk_memcpy_d2h(&r, d_result, sizeof r, nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
    *loss = r.loss;
    last_count = r.count;
    last_wrong = r.wrong;
}

void CrossEntropyLoss::backward() {}

// ------------------------------------------------------------------------------------ ReLU ----
ReLU::ReLU(Variable *in) : in(in) { GCNK_CHECK(gcnk_malloc((void **)&mask, sizeof(uint32_t) * ((size_t)in->size / 32 + 1))); }
ReLU::~ReLU() { gcnk_free(mask); }

void ReLU::forward(bool training) {
    gpu_timer_begin(TMR_RELU_FW);
    GCNK_CHECK(gcnk_relu_fw(in->data, mask, in->size, training, nullptr));
    gpu_timer_end(TMR_RELU_FW);
}

void ReLU::backward() {
    gpu_timer_begin(TMR_RELU_BW);
    GCNK_CHECK(gcnk_relu_bw(in->grad, mask, in->size, nullptr));
    gpu_timer_end(TMR_RELU_BW);
}

// --------------------------------------------------------------------------------- Dropout ----
Dropout::Dropout(Variable *in, float p) : in(in), p(p) {
    GCNK_CHECK(gcnk_malloc((void **)&mask, sizeof(uint32_t) * ((size_t)in->size / 32 + 1)));
}
Dropout::~Dropout() { gcnk_free(mask); }

void Dropout::forward(bool training) {
    if (!training) return;                                   // eval consumes no random numbers (module.cpp:208)
    gpu_timer_begin(TMR_DROPOUT_FW);
    GCNK_CHECK(gcnk_dropout_mask(global_rng(), mask, in->size, p, nullptr));
    GCNK_CHECK(gcnk_dropout_apply(in->data, mask, in->size, p, nullptr));
    gpu_timer_end(TMR_DROPOUT_FW);
}

void Dropout::backward() {
    if (!in->grad) return;                                   // the input features carry no gradient (module.cpp:223-224)
    gpu_timer_begin(TMR_DROPOUT_BW);
    GCNK_CHECK(gcnk_dropout_apply(in->grad, mask, in->size, p, nullptr));
    gpu_timer_end(TMR_DROPOUT_BW);
}
