// optim.h — the Adam optimiser.
//
// Call-compatible with the reference's optimiser (AdamParams / Adam with a list of (Variable*, decay) pairs and
// step(): reference src/seq/optim.h:6-27, GPU twin CUDAAdam in src/cuda/cuda_module.cuh:96-105) and numerically the
// same update (optim.cpp:24-37: L2 folded into the gradient of the decayed variables, fp32 step size from powf/sqrtf,
// the (1 - beta) terms in double).  Here all variables are updated by ONE multi-tensor kernel launch, and the first
// and second moments start at zero (the reference GPU engine leaves them uninitialised, cuda_module.cu:229-233).
#pragma once
#include <utility>
#include <vector>

#include "variable.h"

struct AdamParams {
    float lr, beta1, beta2, eps, weight_decay;
    static AdamParams get_default();          // {0.001, 0.9, 0.999, 1e-8, 0}
};

// One optimised tensor: the Variable (not owned) and its two moment buffers (owned, device).
struct AdamVariable {
    AdamVariable(Variable *var, bool decay);
    AdamVariable(AdamVariable &&other) noexcept;
    AdamVariable(const AdamVariable &) = delete;
    ~AdamVariable();
    int size() const { return var->size; }

    Variable *var;
    float *m = nullptr, *v = nullptr;
    bool decay;
};

class Adam {
public:
    Adam() {}
    Adam(std::vector<std::pair<Variable *, bool>> vars, AdamParams params);
    Adam(Adam &&) = default;
    Adam &operator=(Adam &&other) noexcept;

    // One update of every variable.  d_sumsq (optional, device float) receives sum(w^2) of the FIRST variable after the
    // update: the L2 penalty of the next pass (gcn.cpp:98-105) without another reduction launch.
    void step(float *d_sumsq = nullptr, void *stream = nullptr);   // stream: a gcnk_stream_t (NULL = the legacy stream)

private:
    AdamParams params;
    int step_count = 0;
    std::vector<AdamVariable> vars;
};
