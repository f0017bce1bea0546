#include "optim.h"

#include <cmath>

#include "check.h"
#include "timer.h"

AdamParams AdamParams::get_default() { return {0.001f, 0.9f, 0.999f, 1e-8f, 0.0f}; }   // optim.cpp:6-8

AdamVariable::AdamVariable(Variable *var_, bool decay_) : var(var_), decay(decay_) {
    const size_t bytes = sizeof(float) * (size_t)var->size;
    GCNK_CHECK(gcnk_malloc((void **)&m, bytes));
    GCNK_CHECK(gcnk_malloc((void **)&v, bytes));
    GCNK_CHECK(gcnk_memset(m, 0, bytes, nullptr));
    GCNK_CHECK(gcnk_memset(v, 0, bytes, nullptr));
}
AdamVariable::~AdamVariable() { if (m) gcnk_free(m); if (v) gcnk_free(v); }
AdamVariable::AdamVariable(AdamVariable &&o) noexcept : var(o.var), m(o.m), v(o.v), decay(o.decay) { o.m = o.v = nullptr; }

Adam::Adam(std::vector<std::pair<Variable *, bool>> vars_, AdamParams params_) : params(params_) {
    vars.reserve(vars_.size());
    for (auto &pr : vars_) vars.emplace_back(pr.first, pr.second);
}

Adam &Adam::operator=(Adam &&o) noexcept {
    params = o.params; step_count = o.step_count;
    vars.clear();
    vars.reserve(o.vars.size());
    for (auto &v : o.vars) vars.emplace_back(std::move(v));
    o.vars.clear();
    return *this;
}

void Adam::step(float *d_sumsq, void *stream) {
    step_count++;
    // fp32 powf / sqrtf exactly as optim.cpp:26
    const float step_size = params.lr * sqrtf(1 - powf(params.beta2, step_count)) / (1 - powf(params.beta1, step_count));
    std::vector<gcnk_adam_tensor> t(vars.size());
    for (size_t i = 0; i < vars.size(); i++)
        t[i] = gcnk_adam_tensor{vars[i].var->data, vars[i].var->grad, vars[i].m, vars[i].v, vars[i].size(), vars[i].decay ? 1 : 0};
    gpu_timer_begin(TMR_ADAM);
    GCNK_CHECK(gcnk_adam_step(t.data(), (int)t.size(), step_size, params.beta1, params.beta2, params.eps, params.weight_decay,
                              d_sumsq, stream));
    gpu_timer_end(TMR_ADAM);
}
