// matmul_tc.cu — Matmul forward on the 5th-generation tensor cores: C[M x N] = A[M x K] * B[K x N], fp32 in and out,
// for the wide-hidden layer-2 product of the reference (Matmul::forward, module.cpp:11-22, at hidden 256 x 47 classes:
// the ogbn-products-shape config).  At hidden 16 the product is fused into layer2.cu and never reaches this file.
//
// Structure (one CTA per SM, persistent over 128-row tiles of A; 12 warps):
//   warp 0 (one lane)   TMA producer: cp.async.bulk.tensor 2D loads of A tiles [128 rows x 32 floats] into a 3-stage
//                       ring, 128-byte swizzle, completion on an mbarrier (A's row pitch K*4 bytes is a multiple of 16)
//   warps 4-7           splitters: fp32 has 24 significant bits, a TF32 operand 11, so every A tile is split IN SHARED
//                       MEMORY into big = trunc_tf32(a) (in place) and small = trunc_tf32(a - big) (second buffer); the
//                       split is elementwise, so the swizzled layout is preserved without any address arithmetic
//   warp 1 (one lane)   MMA issuer: tcgen05.mma kind::tf32, M = 128, N = padded class count, K = 8 per instruction,
//                       three per k-step (small*big, big*small, big*big — 3xTF32, fp32-accurate products), accumulator
//                       in TMEM; tcgen05.commit releases the ring slot and finally signals the epilogue
//   warp 2              allocates / frees the TMEM columns
//   warps 8-11          epilogue: tcgen05.ld 32 lanes x 32 bit x 16 columns per warp quadrant -> registers -> C; two
//                       accumulators in TMEM, so the MMAs of the next tile run under the epilogue of this one
// B (the weights: tiny) is transposed, padded and split once per call by prep_b_kernel into K-major [Npad x Kpad]
// big / small copies and stays resident in shared memory for the whole kernel (loaded by TMA with the same swizzle).
//
// Every mbarrier wait has a clock-based bail-out that raises an error flag instead of hanging the GPU.
#include <stdlib.h>

#include <algorithm>

#include "tma.cuh"

using namespace gcnk;

namespace {

constexpr int TC_THREADS = 384;                        // warps 0-3: TMA / MMA / TMEM alloc / idle, 4-7: splitters, 8-11: epilogue
constexpr int BM = 128, BK = 32, STAGES = 3, ACC_STAGES = 2, TMEM_COLS = 128;   // two 64-column accumulators
constexpr int EPI_STRIDE = 65;                         // floats per staged row (Npad <= 64; odd => conflict-free column access)
constexpr int A_TILE_BYTES = BM * BK * 4;              // 16 KB
constexpr uint32_t TF32_MASK = 0xffffe000u;

// ------------------------------------------------------------------------- tcgen05 PTX wrappers ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row atoms 1024 bytes apart (SBO), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

struct Bars {
    uint64_t full[STAGES], ready[STAGES], empty[STAGES], b_full, acc_full[ACC_STAGES], acc_empty[ACC_STAGES];
    uint32_t tmem_base;
};

// B[K x N] row-major  ->  Bt_big / Bt_small [Npad x Kpad], K contiguous, zero padded, split into two TF32 values
__global__ void prep_b_kernel(const float *__restrict__ b, float *__restrict__ bt_big, float *__restrict__ bt_small, int K, int N,
                              int Kpad, int Npad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kpad * Npad) return;
    const int n = i / Kpad, k = i % Kpad;
    const float v = (n < N && k < K) ? b[(size_t)k * N + n] : 0.f;
    const uint32_t big = __float_as_uint(v) & TF32_MASK;
    bt_big[i] = __uint_as_float(big);
    bt_small[i] = __uint_as_float(__float_as_uint(v - __uint_as_float(big)) & TF32_MASK);
}

__global__ void __launch_bounds__(TC_THREADS, 1) matmul_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_bb,
                                                                   const __grid_constant__ CUtensorMap map_bs, float *__restrict__ c,
                                                                   int M, int N, int Npad, int kblocks, int *err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer (LDS, not LD)   // swizzle atoms are 1 KB
    // [A big: STAGES x 16 KB][A small: STAGES x 16 KB][B big: kblocks x Npad x 128 B][B small: same][barriers]
    uint8_t *a_big = smem, *a_small = smem + STAGES * A_TILE_BYTES;
    const uint32_t b_block_bytes = (uint32_t)Npad * BK * 4;
    uint8_t *b_big = a_small + STAGES * A_TILE_BYTES, *b_small = b_big + (size_t)kblocks * b_block_bytes;
    float *epi = reinterpret_cast<float *>(b_small + (size_t)kblocks * b_block_bytes);        // [4 warps][32 rows][EPI_STRIDE]
    Bars *bars = reinterpret_cast<Bars *>(epi + 4 * 32 * EPI_STRIDE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (M + BM - 1) / BM;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->ready[s], 4); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->b_full, 1);
        for (int i = 0; i < ACC_STAGES; i++) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 4); }
        mbar_fence_init();
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0 && lane == 0) {
        // ================================ TMA producer ================================
        mbar_expect_tx(&bars->b_full, 2u * kblocks * b_block_bytes);
        for (int kb = 0; kb < kblocks; kb++) {
            tma_load_2d(b_big + (size_t)kb * b_block_bytes, &map_bb, &bars->b_full, kb * BK, 0);
            tma_load_2d(b_small + (size_t)kb * b_block_bytes, &map_bs, &bars->b_full, kb * BK, 0);
        }
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kb = 0; kb < kblocks; kb++, it++) {
                const int s = it % STAGES;
                if (it >= STAGES && !mbar_wait(&bars->empty[s], ((it / STAGES) - 1) & 1, err)) goto teardown;
                mbar_expect_tx(&bars->full[s], A_TILE_BYTES);
                tma_load_2d(a_big + s * A_TILE_BYTES, &map_a, &bars->full[s], kb * BK, tile * BM);
            }
    } else if (warp == 1 && lane == 0) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Npad >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        if (!mbar_wait(&bars->b_full, 0, err)) goto teardown;
        uint32_t it = 0, t = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t++) {
            const uint32_t acc = t % ACC_STAGES, d_tmem = tmem + acc * 64;
            if (t >= ACC_STAGES && !mbar_wait(&bars->acc_empty[acc], ((t / ACC_STAGES) - 1) & 1, err)) goto teardown;   // epilogue drained it
            tc_fence_after();
            for (int kb = 0; kb < kblocks; kb++, it++) {
                const int s = it % STAGES;
                if (!mbar_wait(&bars->ready[s], (it / STAGES) & 1, err)) goto teardown;
                tc_fence_after();
                const uint32_t ab = smem_u32(a_big + s * A_TILE_BYTES), as = smem_u32(a_small + s * A_TILE_BYTES);
                const uint32_t bb = smem_u32(b_big + (size_t)kb * b_block_bytes), bs = smem_u32(b_small + (size_t)kb * b_block_bytes);
#pragma unroll
                for (int k = 0; k < BK / 8; k++) {
                    const uint32_t off = k * 32;                      // 8 floats inside the 128-byte swizzle atom
                    tc_mma_tf32(d_tmem, make_desc(as + off), make_desc(bb + off), idesc, (kb | k) != 0);
                    tc_mma_tf32(d_tmem, make_desc(ab + off), make_desc(bs + off), idesc, 1);
                    tc_mma_tf32(d_tmem, make_desc(ab + off), make_desc(bb + off), idesc, 1);
                }
                tc_commit(&bars->empty[s]);                           // the slot is free once these MMAs have read it
            }
            tc_commit(&bars->acc_full[acc]);
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================ splitters ================================
        const int tid = threadIdx.x - 128;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kb = 0; kb < kblocks; kb++, it++) {
                const int s = it % STAGES;
                if (!mbar_wait(&bars->full[s], (it / STAGES) & 1, err)) goto teardown;
                float4 *big = reinterpret_cast<float4 *>(a_big + s * A_TILE_BYTES), *sm = reinterpret_cast<float4 *>(a_small + s * A_TILE_BYTES);
#pragma unroll
                for (int i = 0; i < A_TILE_BYTES / 16 / 128; i++) {
                    const float4 v = big[tid + 128 * i];
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(v.x) & TF32_MASK); lo.x = __uint_as_float(__float_as_uint(v.x - hi.x) & TF32_MASK);
                    hi.y = __uint_as_float(__float_as_uint(v.y) & TF32_MASK); lo.y = __uint_as_float(__float_as_uint(v.y - hi.y) & TF32_MASK);
                    hi.z = __uint_as_float(__float_as_uint(v.z) & TF32_MASK); lo.z = __uint_as_float(__float_as_uint(v.z - hi.z) & TF32_MASK);
                    hi.w = __uint_as_float(__float_as_uint(v.w) & TF32_MASK); lo.w = __uint_as_float(__float_as_uint(v.w - hi.w) & TF32_MASK);
                    big[tid + 128 * i] = hi;
                    sm[tid + 128 * i] = lo;
                }
                fence_proxy_async();                                  // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->ready[s]);
            }
    } else if (warp >= 8) {
        // ======================= epilogue (TMEM lane quadrant = warp % 4) =======================
        const int q = warp - 8;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t++) {
            const uint32_t acc = t % ACC_STAGES;
            if (!mbar_wait(&bars->acc_full[acc], (t / ACC_STAGES) & 1, err)) goto teardown;
            tc_fence_after();
            // TMEM -> registers -> this warp's staging rows; then the warp's 32 output rows, which are one contiguous
            // block of 32*N floats in C, go out with coalesced stores
            float *stage = epi + q * 32 * EPI_STRIDE;
            for (int c0 = 0; c0 < Npad; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem + acc * 64 + ((uint32_t)(q * 32) << 16) + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; j++) stage[lane * EPI_STRIDE + c0 + j] = __uint_as_float(r[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);          // the accumulator is free as soon as it is in registers/smem
            {
                const int row0 = tile * BM + q * 32;
                const int rows = min(32, M - row0);
                if (rows > 0) {
                    float *out = c + (size_t)row0 * N;
                    const int total = rows * N;
                    for (int f = lane; f < total; f += 32) {
                        const int rr = f / N, cc = f - rr * N;
                        out[f] = stage[rr * EPI_STRIDE + cc];
                    }
                }
            }
            __syncwarp();                                               // the staging rows are reused by the next tile
        }
    }
teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

}  // namespace

namespace gcnk {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
static EncodeTiled encode_fn() {
    static EncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiled>(p);
    }
    return fn;
}
bool tensor_maps_available() { return encode_fn() != nullptr; }

bool make_tensor_map_2d(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint64_t pitch_floats, uint32_t box_rows,
                        uint32_t box_cols, bool swizzle128) {
    EncodeTiled fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {pitch_floats * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shapes this kernel takes: enough rows to matter, A's row pitch a multiple of 16 bytes, both split copies of B resident
bool matmul_tc_supported(int m, int k, int n) {
    static const bool off = getenv("GCNK_NO_TCGEN05") && atoi(getenv("GCNK_NO_TCGEN05")) != 0;
    if (off || m < 1024 || k < 64 || k % 4 || n < 8 || n > 256) return false;
    const int npad = (n + 15) / 16 * 16, kpad = (k + BK - 1) / BK * BK;
    const size_t smem = 2 * (size_t)STAGES * A_TILE_BYTES + 2 * (size_t)npad * kpad * 4 + 4 * 32 * EPI_STRIDE * sizeof(float) + sizeof(Bars) + 1024;
    return npad <= 64 && smem <= 227 * 1024 && tensor_maps_available();
}

int matmul_tc_fw(const float *a, const float *b, float *c, int m, int k, int n, cudaStream_t st) {
    const int npad = (n + 15) / 16 * 16, kpad = (k + BK - 1) / BK * BK, kblocks = kpad / BK;
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    static float *bt[64] = {nullptr};
    static size_t bt_elems[64] = {0};
    const size_t need = 2 * (size_t)npad * kpad;
    if (bt_elems[dev] < need) {
        GCNK_CUDA(cudaStreamSynchronize(st));
        if (bt[dev]) GCNK_CUDA(cudaFree(bt[dev]));
        GCNK_CUDA(cudaMalloc(&bt[dev], sizeof(float) * need));
        bt_elems[dev] = need;
    }
    float *bt_big = bt[dev], *bt_small = bt[dev] + (size_t)npad * kpad;
    prep_b_kernel<<<(npad * kpad + 255) / 256, 256, 0, st>>>(b, bt_big, bt_small, k, n, kpad, npad);
    GCNK_LAUNCHED();
    CUtensorMap map_a, map_bb, map_bs;
    if (!make_tensor_map_2d(&map_a, a, (uint64_t)m, (uint64_t)k, (uint64_t)k, BM, BK, true) ||
        !make_tensor_map_2d(&map_bb, bt_big, (uint64_t)npad, (uint64_t)kpad, (uint64_t)kpad, (uint32_t)npad, BK, true) ||
        !make_tensor_map_2d(&map_bs, bt_small, (uint64_t)npad, (uint64_t)kpad, (uint64_t)kpad, (uint32_t)npad, BK, true)) {
        set_error("matmul_tc: cuTensorMapEncodeTiled failed");
        return GCNK_EUNSUPPORTED;
    }
    const size_t smem = 2 * (size_t)STAGES * A_TILE_BYTES + 2 * (size_t)npad * kpad * 4 + 4 * 32 * EPI_STRIDE * sizeof(float) + sizeof(Bars) + 1024;
    static bool attr[64] = {false};
    if (!attr[dev]) {
        GCNK_CUDA(cudaFuncSetAttribute(matmul_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr[dev] = true;
    }
    const int n_tiles = (m + BM - 1) / BM;
    matmul_tc_kernel<<<std::min(n_tiles, sm_count()), TC_THREADS, smem, st>>>(map_a, map_bb, map_bs, c, m, n, npad, kblocks, async_err_flag());
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // namespace gcnk
