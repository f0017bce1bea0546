// optim.h — Adam with the reference's public shape (src/seq/optim.h:6-27; GPU twin CUDAAdam,
// src/cuda/cuda_module.cuh:96-105).  One fused multi-tensor launch per step; m and v start at zero
// (the reference GPU path leaves them uninitialised, cuda_module.cu:229-233).
#pragma once
#include <utility>
#include <vector>

#include "variable.h"

struct AdamParams {
    float lr, beta1, beta2, eps, weight_decay;
    static AdamParams get_default();
};

struct AdamVariable {
    Variable *var;
    float *m = nullptr, *v = nullptr;   // device
    bool decay;
    int size() const { return var->size; }
    AdamVariable(Variable *var, bool decay);
    ~AdamVariable();
    AdamVariable(AdamVariable &&o) noexcept;
    AdamVariable(const AdamVariable &) = delete;
};

class Adam {
    AdamParams params;
    int step_count = 0;
    std::vector<AdamVariable> vars;
public:
    Adam() {}
    Adam(std::vector<std::pair<Variable *, bool>> vars, AdamParams params);
    Adam(Adam &&) = default;
    Adam &operator=(Adam &&o) noexcept;
    // d_sumsq (optional, device float): receives sum(w^2) of the first variable after the update
    void step(float *d_sumsq = nullptr);
};
