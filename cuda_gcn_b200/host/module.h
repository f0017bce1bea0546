// module.h — the operator API of the reference (src/seq/module.h:6-76; GPU twin
// src/cuda/cuda_module.cuh:11-94): abstract Module{forward(bool training), backward()} and the six
// operators with the reference's constructor signatures.  Each operator is individually callable and
// binds to the corresponding UNFUSED entry point of the C ABI (include/gcnk.h); the GCN driver's
// fused plan (gcn.cpp) calls the fused entry points directly instead.  There is no CPU path.
#pragma once
#include <cstdint>

#include "sparse.h"
#include "variable.h"

class Module {
public:
    virtual void forward(bool training) = 0;
    virtual void backward() = 0;
    virtual ~Module() {}
};

// c[m x p] = a[m x n] * b[n x p]                                  (module.cpp:11-42)
class Matmul : public Module {
    Variable *a, *b, *c;
    int m, n, p;
    float *workspace = nullptr;
    size_t workspace_bytes = 0;
public:
    Matmul(Variable *a, Variable *b, Variable *c, int m, int n, int p);
    ~Matmul();
    void forward(bool);
    void backward();
};

// c[m x p] = CSR(sp, values = a->data)[m x n] * b[n x p]; only b gets a gradient   (module.cpp:47-77)
class SparseMatmul : public Module {
    Variable *a, *b, *c;
    SparseIndex *sp;
    int m, n, p;
public:
    SparseMatmul(Variable *a, Variable *b, Variable *c, SparseIndex *sp, int m, int n, int p);
    ~SparseMatmul() {}
    void forward(bool);
    void backward();
};

// out = A_hat * in; backward: in->grad = A_hat * out->grad (the same product, module.cpp:83-119)
class GraphSum : public Module {
    Variable *in, *out;
    SparseIndex *graph;
    int dim;
public:
    GraphSum(Variable *in, Variable *out, SparseIndex *graph, int dim);
    ~GraphSum() {}
    void forward(bool);
    void backward();
};

// truth is a DEVICE int[n] as in the reference's GPU twin (cuda_gcn.cu:58-59); *loss is a host float.
// forward also counts the wrongly classified labelled rows (GCN::get_accuracy's rule, gcn.cpp:83-96).
class CrossEntropyLoss : public Module {
    Variable *logits;
    int *truth;
    float *loss;
    int num_classes;
    gcnk_ce_result *d_result = nullptr;
    float *workspace = nullptr;
    size_t workspace_bytes = 0;
public:
    int last_count = 0, last_wrong = 0;
    CrossEntropyLoss(Variable *logits, int *truth, float *loss, int num_classes);
    ~CrossEntropyLoss();
    void forward(bool);
    void backward();
};

class ReLU : public Module {
    Variable *in;
    uint32_t *mask;      // 1 bit per element (the reference keeps a bool per element, module.cpp:166-173)
public:
    ReLU(Variable *in);
    ~ReLU();
    void forward(bool);
    void backward();
};

// keep = (int)RAND() >= int(p * MY_RAND_MAX) drawn from the process-wide xorshift128+ stream in
// element order, exactly as gcn-seq (module.cpp:207-221) — on the device, in parallel.
class Dropout : public Module {
    Variable *in;
    uint32_t *mask;      // keep bits, 1 per element; kept even when in->grad == nullptr (needed for the values)
    float p;
public:
    Dropout(Variable *in, float p);
    ~Dropout();
    void forward(bool);
    void backward();
};
