// common.cuh — shared plumbing for libgcnk.so (error capture, launch accounting, small device helpers).
#pragma once
#include <cstdlib>
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

// Shared-memory carve-out of the kernels that run CONCURRENTLY on different streams (the gathers, the keep-bit generator
// on its low-priority stream, the sequential loss sum): an SM can only switch its L1 / shared-memory split when it is
// empty, so kernels that ask for different splits cannot share an SM and the block scheduler drains SMs to make room.
// GCN_CARVEOUT=<percent> gives all of them the same preference (-1 / unset: the driver's per-kernel choice).
inline int carveout_pct() {
    static const int v = [] { const char *e = getenv("GCN_CARVEOUT"); return e && *e ? atoi(e) : -1; }();
    return v;
}
template <typename Kernel>
inline void prefer_carveout(Kernel k) {
    const int pct = carveout_pct();
    if (pct >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
}
long long peer_spin_cycles();   // clock64 ticks a kernel waits for a peer's flag before raising its error flag (GCN_PEER_TIMEOUT_S)
int *async_err_flag();   // per-device int raised by pipeline kernels whose mbarrier wait timed out

// Mirrored output rows (row-partitioned runs): a producer kernel whose output is the input of the next GraphSum
// on EVERY rank stores each row it writes also at the same offset of the peers' buffers, over NVLink peer
// mappings — the all-gather is fused into the producer's epilogue instead of being a separate collective.
constexpr int MAX_PEERS = 7;
struct Mirror {
    float *p[MAX_PEERS];
    int n;
};
// the mirror registered with gcnk_mirror_next for `out` (n == 0 if none); consumes the registration
Mirror take_mirror(const float *out);

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

__device__ __forceinline__ void mirror_store(const Mirror &m, size_t idx, float4 v) {
    for (int i = 0; i < m.n; i++) *reinterpret_cast<float4 *>(m.p[i] + idx) = v;
}
__device__ __forceinline__ void mirror_store(const Mirror &m, size_t idx, float2 v) {
    for (int i = 0; i < m.n; i++) *reinterpret_cast<float2 *>(m.p[i] + idx) = v;
}
__device__ __forceinline__ void mirror_store(const Mirror &m, size_t idx, float v) {
    for (int i = 0; i < m.n; i++) m.p[i][idx] = v;
}

// out[i] = sum over b < parts of partials[b*elems + i], deterministically: thread (e, q) of a 32 x 8 CTA sums the
// parts b = q, q+8, ... of element e in order (two independent chains), the eight group sums are then added in
// group order.  Replaces a single serial chain of `parts` dependent L2 reads per element (44 us for 592 parts).
__device__ __forceinline__ void reduce_parts_block(const float *__restrict__ partials, float *__restrict__ out, int elems, int parts) {
    __shared__ float s_part[8][33];
    const int e = threadIdx.x & 31, q = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + e;
    // eight independent chains per thread: the loop is pure load latency (measured: 45 us for 911 partials with two)
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (i < elems) {
        int b = q;
        for (; b + 56 < parts; b += 64) {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] += partials[(size_t)(b + 8 * k) * elems + i];
        }
        for (int k = 0; b < parts; b += 8, k++) s[k] += partials[(size_t)b * elems + i];
    }
    s_part[q][e] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    __syncthreads();
    if (q == 0 && i < elems) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) v += s_part[k][e];
        out[i] = v;
    }
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
