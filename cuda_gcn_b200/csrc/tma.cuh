// tma.cuh — thin PTX wrappers for the Blackwell async machinery used by matmul_tc.cu and feature_tma.cu:
// mbarriers, 2-D TMA tensor loads (cp.async.bulk.tensor), proxy fences, and the host-side tensor-map encoder
// (cuTensorMapEncodeTiled through the runtime's driver entry point, so nothing links against libcuda).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gcnk {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// returns false on timeout (~1 s) so that a protocol error cannot hang the device; *err is set to 2
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *err) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
        if (clock64() - t0 > (2LL << 30)) { *err = 2; return false; }
    }
}
// box of a 2-D tensor -> shared memory; completion (bytes) is signalled on `bar`.  c0 = innermost coordinate.
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of element (row r, float column c < 32) inside a [rows x 32 floats] box written by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B into a 1 KB-aligned buffer: the 16-byte chunk index is XORed with (r & 7)
__device__ __forceinline__ uint32_t sw128_offset(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2))); }

// 2-D fp32 tensor [rows x cols] with row pitch `pitch_floats` (pitch*4 must be a multiple of 16), box [box_rows x box_cols],
// zero fill out of bounds.  swizzle128: box_cols must be 32 (128-byte rows).
bool make_tensor_map_2d(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint64_t pitch_floats, uint32_t box_rows,
                        uint32_t box_cols, bool swizzle128);
bool tensor_maps_available();

}  // namespace gcnk
