// variable.h — Variable: a {data, grad} pair of fp32 DEVICE buffers with the reference's public
// shape (src/seq/variable.h:4-12; GPU twin src/cuda/cuda_variable.cuh:7-20): same constructor,
// glorot / zero / zero_grad / print / grad_norm.  Buffers are owned (RAII); Modules hold raw
// non-owning pointers to Variables exactly as in the reference (module.h:14,24-25,35-36).
#pragma once
#include <vector>

struct Variable {
    float *data = nullptr, *grad = nullptr;   // device pointers; grad == nullptr when !requires_grad
    int size = 0;

    Variable(int size, bool requires_grad = true);
    ~Variable();
    Variable(Variable &&o) noexcept;
    Variable &operator=(Variable &&o) noexcept;
    Variable(const Variable &) = delete;
    Variable &operator=(const Variable &) = delete;

    // Glorot-uniform from the process-wide xorshift128+ stream, drawn on the host in the reference's
    // order and arithmetic (variable.cpp:11-18) so that weights are bit-identical to gcn-seq's for the
    // same seed, then uploaded (the reference GPU path uses an unrelated cuRAND stream, cuda_kernel.cu:290-295).
    void glorot(int in_size, int out_size);
    void zero();
    void zero_grad();
    void print(int col = 0x7fffffff);
    float grad_norm();

    // host <-> device helpers (synchronous); not part of the reference API
    std::vector<float> host_data() const;
    std::vector<float> host_grad() const;
    void set_data(const float *h_src);
    void set_grad(const float *h_src);
};
