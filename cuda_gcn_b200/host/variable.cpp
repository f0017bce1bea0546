#include "variable.h"

#include <cmath>
#include <cstdio>

#include "check.h"
#include "rand.h"

Variable::Variable(int size_, bool requires_grad) : size(size_) {
    GCNK_CHECK(gcnk_malloc((void **)&data, sizeof(float) * (size_t)size));
    if (requires_grad) GCNK_CHECK(gcnk_malloc((void **)&grad, sizeof(float) * (size_t)size));
    // std::vector<float>(size) value-initialises in the reference; kernels here write their outputs
    // fully, but a freshly constructed Variable still reads as zeros
    zero();
    zero_grad();
}

Variable::~Variable() {
    if (data) gcnk_free(data);
    if (grad) gcnk_free(grad);
}

Variable::Variable(Variable &&o) noexcept : data(o.data), grad(o.grad), size(o.size) { o.data = o.grad = nullptr; o.size = 0; }

Variable &Variable::operator=(Variable &&o) noexcept {
    if (this != &o) {
        if (data) gcnk_free(data);
        if (grad) gcnk_free(grad);
        data = o.data; grad = o.grad; size = o.size;
        o.data = o.grad = nullptr; o.size = 0;
    }
    return *this;
}

void Variable::glorot(int in_size, int out_size) {
    const float range = sqrtf(6.0f / (in_size + out_size));
    std::vector<uint32_t> draws((size_t)size);
    GCNK_CHECK(gcnk_rng_next_host(global_rng(), draws.data(), size));
    std::vector<float> w((size_t)size);
    for (int i = 0; i < size; i++) {
        // float(RAND()) / MY_RAND_MAX  (an int, converted to float 2^31), minus 0.5 in double, narrowed
        const float r = (float)((double)((float)draws[i] / (float)MY_RAND_MAX) - 0.5);
        w[i] = r * range * 2;
    }
    set_data(w.data());
}

void Variable::zero() { GCNK_CHECK(gcnk_memset(data, 0, sizeof(float) * (size_t)size, nullptr)); }
void Variable::zero_grad() { if (grad) GCNK_CHECK(gcnk_memset(grad, 0, sizeof(float) * (size_t)size, nullptr)); }

std::vector<float> Variable::host_data() const {
    std::vector<float> h((size_t)size);
    GCNK_CHECK(gcnk_memcpy_d2h(h.data(), data, sizeof(float) * h.size(), nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
    return h;
}

std::vector<float> Variable::host_grad() const {
    std::vector<float> h((size_t)(grad ? size : 0));
    if (grad) {
        GCNK_CHECK(gcnk_memcpy_d2h(h.data(), grad, sizeof(float) * h.size(), nullptr));
        GCNK_CHECK(gcnk_stream_sync(nullptr));
    }
    return h;
}

void Variable::set_data(const float *h) {
    GCNK_CHECK(gcnk_memcpy_h2d(data, h, sizeof(float) * (size_t)size, nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
}

void Variable::set_grad(const float *h) {
    if (!grad) return;
    GCNK_CHECK(gcnk_memcpy_h2d(grad, h, sizeof(float) * (size_t)size, nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
}

void Variable::print(int col) {
    const std::vector<float> h = host_data();
    int count = 0;
    for (float x : h) {
        printf("%.4f ", x);
        if (++count % col == 0) printf("\n");
    }
    printf("\n");
}

float Variable::grad_norm() {
    float norm = 0;
    for (float x : host_grad()) norm += x * x;
    return sqrtf(norm);
}
