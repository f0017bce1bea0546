// module.h — the operator interface of this engine.
//
// It is deliberately call-compatible with the reference's operator layer (abstract `Module` with
// `forward(bool training)` / `backward()`, reference src/seq/module.h:6-11 and its GPU twin
// src/cuda/cuda_module.cuh:11-16) and with the constructor argument lists of its six operators
// (module.h:17,28,39,51,61,72), so that code written against the reference keeps compiling.  What is
// behind the interface is different: every operator owns only device-side scratch, reads and writes
// the DEVICE buffers of the `Variable`s it was given (non-owning pointers, as in the reference), and
// binds to one UNFUSED entry point of the kernel C ABI (include/gcnk.h).  The GCN driver's fused plan
// (gcn.cpp) bypasses these objects and calls the fused entry points directly.  There is no CPU path.
#pragma once
#include <cstddef>
#include <cstdint>

#include "sparse.h"
#include "variable.h"

class Module {
public:
    virtual ~Module() {}
    virtual void forward(bool training) = 0;
    virtual void backward() = 0;
};

// ---- dense product: c[m x p] = a[m x n] * b[n x p]                                   (module.cpp:11-42)
//      backward: a->grad = c->grad * b^T ; b->grad = a^T * c->grad (split over rows, fixed-order reduction)
class Matmul final : public Module {
public:
    Matmul(Variable *a, Variable *b, Variable *c, int m, int n, int p);
    ~Matmul() override;
    void forward(bool training) override;
    void backward() override;

private:
    Variable *const a, *const b, *const c;
    const int m, n, p;
    float *workspace = nullptr;          // split-K partials of the b gradient
    size_t workspace_bytes = 0;
};

// ---- sparse feature transform: c[m x p] = CSR(sp; values = a->data)[m x n] * b[n x p]   (module.cpp:47-77)
//      only b receives a gradient (the features are inputs)
class SparseMatmul final : public Module {
public:
    SparseMatmul(Variable *a, Variable *b, Variable *c, SparseIndex *sp, int m, int n, int p);
    void forward(bool training) override;
    void backward() override;

private:
    Variable *const a, *const b, *const c;
    SparseIndex *const sp;
    const int m, n, p;
};

// ---- normalised-adjacency aggregation: out = A_hat * in                               (module.cpp:83-119)
//      backward is the same product on the gradients (in->grad = A_hat * out->grad), exactly as the reference
class GraphSum final : public Module {
public:
    GraphSum(Variable *in, Variable *out, SparseIndex *graph, int dim);
    void forward(bool training) override;
    void backward() override;

private:
    Variable *const in, *const out;
    SparseIndex *const graph;
    const int dim;
};

// ---- masked softmax cross-entropy                                                     (module.cpp:124-164)
//      `truth` is a DEVICE int[n] (-1 = row not in the split), as in the reference's GPU engine
//      (cuda_gcn.cu:58-59); `*loss` is a host float written by forward().  The same pass counts the
//      wrongly classified labelled rows with GCN::get_accuracy's rule (gcn.cpp:83-96).
class CrossEntropyLoss final : public Module {
public:
    CrossEntropyLoss(Variable *logits, int *truth, float *loss, int num_classes);
    ~CrossEntropyLoss() override;
    void forward(bool training) override;
    void backward() override;                 // nothing to do: forward already left d(loss)/d(logits) in logits->grad

    int last_count = 0, last_wrong = 0;       // labelled rows / wrongly classified rows of the last forward

private:
    Variable *const logits;
    int *const truth;
    float *const loss;
    const int num_classes;
    gcnk_ce_result *d_result = nullptr;
    float *workspace = nullptr;
    size_t workspace_bytes = 0;
};

// ---- in-place rectifier; the mask is one bit per element (a bool per element in the reference, module.cpp:166-194)
class ReLU final : public Module {
public:
    explicit ReLU(Variable *in);
    ~ReLU() override;
    void forward(bool training) override;     // the mask is only refreshed when training
    void backward() override;

private:
    Variable *const in;
    uint32_t *mask = nullptr;
};

// ---- in-place inverted dropout                                                        (module.cpp:196-233)
//      keep = (int)RAND() >= int(p * MY_RAND_MAX), one draw per element in element order from the process-wide
//      xorshift128+ stream — the same bits gcn-seq draws for the same seed, produced on the device in parallel.
//      The keep bits are stored even when `in` has no gradient (one bit per element instead of the reference's int).
class Dropout final : public Module {
public:
    Dropout(Variable *in, float p);
    ~Dropout() override;
    void forward(bool training) override;     // no-op (and no draws) when !training
    void backward() override;                 // no-op when in->grad == nullptr

private:
    Variable *const in;
    const float p;
    uint32_t *mask = nullptr;
};
