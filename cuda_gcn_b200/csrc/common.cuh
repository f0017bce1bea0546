// common.cuh — shared plumbing for libgcnk.so (error capture, launch accounting, small device helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "gcnk.h"

namespace gcnk {

extern std::atomic<int64_t> g_launches;
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

inline cudaStream_t S(gcnk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define GCNK_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) return gcnk::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// after every <<<>>>: count it and surface launch-configuration errors (the reference checks
// cudaGetLastError after each launch too, e.g. cuda_module.cu:15)
#define GCNK_LAUNCHED()                                                              \
    do {                                                                             \
        gcnk::g_launches.fetch_add(1, std::memory_order_relaxed);                    \
        GCNK_CUDA(cudaGetLastError());                                               \
    } while (0)

#define GCNK_REQUIRE(cond, msg)                                                      \
    do {                                                                             \
        if (!(cond)) { gcnk::set_error("%s: %s", __func__, msg); return GCNK_EINVAL; } \
    } while (0)

int sm_count();          // of the current device (cached per device)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// streaming (read-once) loads: keep them out of L1 so the gather working set stays cached
__device__ __forceinline__ int ld_stream_i32(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

}  // namespace gcnk
