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
//
// Generalised for the wide-hidden plan (host/gcn_wide.cpp) to the three products of a layer (csrc/gemm.cu dispatches):
//   nn / nt  C[m x n] (pitch ldc) = A[m x k] (pitch lda) * B, B given as [k x n] or as [n x k]; n is covered by tiles of
//            at most 64 columns, each CTA keeps ITS tile of B resident and walks the row tiles (the CTAs of one row tile
//            run side by side, so A is re-read from L2, not from HBM); optional row scale in the epilogue.
//   tn       C[ka x n] = A[m x ka]^T * B[m x n]: the node dimension m is the contraction.  Both operands are MN-major
//            for the tensor core (instruction-descriptor bits 15/16; shared-memory descriptors with LBO = the distance of
//            two 32-float column groups, exactly what a [32 rows x 32 floats] TMA box with the 128-byte / 32-byte-atom
//            swizzle lays down — the one layout tcgen05 accepts for MN-major 32-bit operands),
//            so the row-major tiles are fed as they are — no transposed copies.  Split over m across CTAs, partial tiles
//            reduced in a fixed order (deterministic).
// Every mbarrier wait has a clock-based bail-out that raises an error flag instead of hanging the GPU.
#include <stdlib.h>

#include <algorithm>

#include "tma.cuh"

using namespace gcnk;

namespace {

constexpr int TC_THREADS = 384;                        // warps 0-3: TMA / MMA / TMEM alloc / idle, 4-7: splitters, 8-11: epilogue
constexpr int BM = 128, BK = 32, STAGES = 3, ACC_STAGES = 2, TMEM_COLS = 128;   // two 64-column accumulators
// floats per staged epilogue row: Npad + 4 (20, 36, 52, 68) — 16-byte aligned rows, and (Npad + 4) / 4 is odd, which keeps the
// quarter-warp float4 accesses of eight different rows on distinct banks
__host__ __device__ constexpr int epi_stride(int npad) { return npad + 4; }
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
// (b_is_nk: B is already stored [N x K], pitch ldb — the nt product)
__global__ void prep_b_kernel(const float *__restrict__ b, int ldb, int b_is_nk, float *__restrict__ bt_big, float *__restrict__ bt_small, int K, int N,
                              int Kpad, int Npad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kpad * Npad) return;
    const int n = i / Kpad, k = i % Kpad;
    const float v = (n < N && k < K) ? (b_is_nk ? b[(size_t)n * ldb + k] : b[(size_t)k * ldb + n]) : 0.f;
    const uint32_t big = __float_as_uint(v) & TF32_MASK;
    bt_big[i] = __uint_as_float(big);
    bt_small[i] = __uint_as_float(__float_as_uint(v - __uint_as_float(big)) & TF32_MASK);
}

__global__ void __launch_bounds__(TC_THREADS, 1) matmul_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_bb,
                                                                   const __grid_constant__ CUtensorMap map_bs, float *__restrict__ c,
                                                                   int ldc, const float *__restrict__ row_scale, int M, int N, int Npad,
                                                                   int ntn, int kblocks, const uint32_t *__restrict__ keep, int K,
                                                                   float out_scale, int relu, int *err) {
    // keep (optional): Dropout on read — bit (row*K + col) of the reference's flat draw order decides whether A[row, col]
    // takes part (module.cpp:207-224); the 1/(1-p) factor is out_scale, applied with the row scale in the epilogue.
    // Npad = columns of one n tile (a multiple of 16, <= 64); this CTA owns n tile `nt` and walks row tiles
    // tile0, tile0 + tstep, ...
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer (LDS, not LD)   // swizzle atoms are 1 KB
    // [A big: STAGES x 16 KB][A small: STAGES x 16 KB][B big: kblocks x Npad x 128 B][B small: same][barriers]
    uint8_t *a_big = smem, *a_small = smem + STAGES * A_TILE_BYTES;
    const uint32_t b_block_bytes = (uint32_t)Npad * BK * 4;
    uint8_t *b_big = a_small + STAGES * A_TILE_BYTES, *b_small = b_big + (size_t)kblocks * b_block_bytes;
    const int EPI_STRIDE = epi_stride(Npad);
    float *epi = reinterpret_cast<float *>(b_small + (size_t)kblocks * b_block_bytes);        // [4 warps][32 rows][EPI_STRIDE]
    Bars *bars = reinterpret_cast<Bars *>(epi + 4 * 32 * EPI_STRIDE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (M + BM - 1) / BM;
    const int nt = blockIdx.x % ntn, tile0 = blockIdx.x / ntn, tstep = gridDim.x / ntn;   // gridDim.x is a multiple of ntn

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
            tma_load_2d(b_big + (size_t)kb * b_block_bytes, &map_bb, &bars->b_full, kb * BK, nt * Npad);
            tma_load_2d(b_small + (size_t)kb * b_block_bytes, &map_bs, &bars->b_full, kb * BK, nt * Npad);
        }
        uint32_t it = 0;
        for (int tile = tile0; tile < n_tiles; tile += tstep)
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
        for (int tile = tile0; tile < n_tiles; tile += tstep, t++) {
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
        for (int tile = tile0; tile < n_tiles; tile += tstep)
            for (int kb = 0; kb < kblocks; kb++, it++) {
                const int s = it % STAGES;
                if (!mbar_wait(&bars->full[s], (it / STAGES) & 1, err)) goto teardown;
                float4 *big = reinterpret_cast<float4 *>(a_big + s * A_TILE_BYTES), *sm = reinterpret_cast<float4 *>(a_small + s * A_TILE_BYTES);
#pragma unroll
                for (int i = 0; i < A_TILE_BYTES / 16 / 128; i++) {
                    float4 v = big[tid + 128 * i];
                    if (keep) {
                        // float4 `idx` of the swizzled [128 x 32] tile: row idx/8, 16-byte chunk (idx%8) ^ (row%8)
                        const int idx = tid + 128 * i, row = idx >> 3, col = (((idx & 7) ^ (row & 7)) << 2) + kb * BK;
                        const int grow = tile * BM + row;
                        uint32_t nib = 0;
                        if (grow < M && col < K) {
                            const size_t bit = (size_t)grow * K + col;
                            const uint32_t lo32 = keep[bit >> 5], hi32 = (bit & 31) > 28 ? keep[(bit >> 5) + 1] : 0u;
                            nib = __funnelshift_r(lo32, hi32, (uint32_t)(bit & 31)) & 0xfu;
                        }
                        v.x = (nib & 1u) ? v.x : 0.f; v.y = (nib & 2u) ? v.y : 0.f; v.z = (nib & 4u) ? v.z : 0.f; v.w = (nib & 8u) ? v.w : 0.f;
                    }
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
        for (int tile = tile0; tile < n_tiles; tile += tstep, t++) {
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
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4 *>(stage + lane * EPI_STRIDE + c0 + j) =
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);          // the accumulator is free as soon as it is in registers/smem
            {
                const int row0 = tile * BM + q * 32;
                const int rows = min(32, M - row0);
                const int n0 = nt * Npad, cols = min(Npad, N - n0);          // this tile's share of the N real columns
                if (rows > 0 && cols > 0) {
                    float *out = c + (size_t)row0 * ldc + n0;
                    if (((ldc | cols) & 3) == 0 && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
                        // two output rows per step, one float4 per lane: 128-bit shared loads and coalesced 128-bit stores
                        const int c4 = lane & 15, cpr = cols >> 2;
#pragma unroll 4
                        for (int it = 0; it < 16; it++) {
                            const int rr = 2 * it + (lane >> 4);
                            if (c4 < cpr && rr < rows) {
                                float4 v = *reinterpret_cast<const float4 *>(stage + rr * EPI_STRIDE + 4 * c4);
                                if (relu) { v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f; v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f; }
                                const float rs = (row_scale ? row_scale[row0 + rr] : 1.0f) * out_scale;
                                v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
                                *reinterpret_cast<float4 *>(out + (size_t)rr * ldc + 4 * c4) = v;
                            }
                        }
                    } else {
                        const int total = rows * cols;
                        for (int f = lane; f < total; f += 32) {
                            const int rr = f / cols, cc = f - rr * cols;
                            float v = stage[rr * EPI_STRIDE + cc];
                            if (relu) v = v > 0.f ? v : 0.f;
                            out[(size_t)rr * ldc + cc] = (row_scale ? row_scale[row0 + rr] : 1.0f) * out_scale * v;
                        }
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


// ================================================================================== tn kernel ====
// C[ka x n] = A[m x ka]^T * B[m x n], contraction over the rows m.  CTA = (128-feature tile of A, n tile of B, part of m).
// Stage = 32 rows: A as four [32 x 32-float] TMA boxes (one per 32-feature group, 4 KB each, 128-byte swizzle with 32-byte
// atoms) and B as up to four such boxes; in this layout both operands are MN-major for tcgen05 (the contiguous direction
// is M resp. N): 8 rows = two 512-byte swizzle atoms = the K = 8 of one tf32 MMA, column groups LBO = 4096 bytes apart.
constexpr int TN_ROWS = 32, TN_STAGES = 3, TN_BOX_BYTES = TN_ROWS * 128, TN_OPERAND_BYTES = 4 * TN_BOX_BYTES;   // 16 KB
constexpr int TN_STAGE_BYTES = 4 * TN_OPERAND_BYTES;                     // A big, A small, B big, B small

struct TnBars {
    uint64_t full[TN_STAGES], ready[TN_STAGES], empty[TN_STAGES], acc_full;
    uint32_t tmem_base;
};

// MN-major tf32 operand.  The only shared-memory layout the tensor core takes for 32-bit MN-major operands is the
// 128-byte swizzle with a 32-byte base (layout type 1: 32-byte chunks of a 128-byte row XORed with the row index mod 4,
// atoms of 4 rows = 512 bytes) — what a TMA box written with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B looks like.
// Leading-dimension byte offset (between 32-float column groups) = one box = 4096; stride byte offset (between 4-row
// groups along K) = 512; descriptor version 1.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)(TN_BOX_BYTES >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}

__global__ void __launch_bounds__(TC_THREADS, 1) matmul_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                                   float *__restrict__ ws, int ld_ws, int rows_ws, int M, int rows_per_part, int mt,
                                                                   int ntn, int nt_cols, const uint32_t *__restrict__ keep, int KA, int *err) {
    // keep (optional): Dropout on read for the A operand — bit (row*KA + col) of the flat draw order (module.cpp:207-224)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    TnBars *bars = reinterpret_cast<TnBars *>(smem + (size_t)TN_STAGES * TN_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = mt * ntn, tile = blockIdx.x % tiles, part = blockIdx.x / tiles;
    const int m_tile = tile / ntn, n_tile = tile % ntn;
    const int r_lo = part * rows_per_part, r_hi = min(M, r_lo + rows_per_part);
    const int steps = (r_hi - r_lo + TN_ROWS - 1) / TN_ROWS;             // >= 1 by construction of the grid
    const int b_boxes = (nt_cols + 31) / 32;
    (void)mt;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TN_STAGES; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->ready[s], 4); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->acc_full, 1);
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
        for (int it = 0; it < steps; it++) {
            const int s = it % TN_STAGES;
            if (it >= TN_STAGES && !mbar_wait(&bars->empty[s], ((it / TN_STAGES) - 1) & 1, err)) goto teardown;
            uint8_t *st = smem + (size_t)s * TN_STAGE_BYTES;
            mbar_expect_tx(&bars->full[s], (uint32_t)(4 + b_boxes) * TN_BOX_BYTES);
            const int row = r_lo + it * TN_ROWS;
            for (int fg = 0; fg < 4; fg++) tma_load_2d(st + fg * TN_BOX_BYTES, &map_a, &bars->full[s], m_tile * 128 + fg * 32, row);
            for (int fb = 0; fb < b_boxes; fb++)
                tma_load_2d(st + 2 * TN_OPERAND_BYTES + fb * TN_BOX_BYTES, &map_b, &bars->full[s], n_tile * nt_cols + fb * 32, row);
        }
    } else if (warp == 1 && lane == 0) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(nt_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int it = 0; it < steps; it++) {
            const int s = it % TN_STAGES;
            if (!mbar_wait(&bars->ready[s], (it / TN_STAGES) & 1, err)) goto teardown;
            tc_fence_after();
            const uint32_t ab = smem_u32(smem + (size_t)s * TN_STAGE_BYTES), as = ab + TN_OPERAND_BYTES, bb = ab + 2 * TN_OPERAND_BYTES, bs = ab + 3 * TN_OPERAND_BYTES;
#pragma unroll
            for (int k = 0; k < TN_ROWS / 8; k++) {
                const uint32_t off = k * 1024;                            // the next 8 rows = the next swizzle atom
                tc_mma_tf32(tmem, make_desc_mn(as + off), make_desc_mn(bb + off), idesc, (it | k) != 0);
                tc_mma_tf32(tmem, make_desc_mn(ab + off), make_desc_mn(bs + off), idesc, 1);
                tc_mma_tf32(tmem, make_desc_mn(ab + off), make_desc_mn(bb + off), idesc, 1);
            }
            tc_commit(&bars->empty[s]);
        }
        tc_commit(&bars->acc_full);
    } else if (warp >= 4 && warp < 8) {
        // ================================ splitters (both operands) ================================
        const int tid = threadIdx.x - 128;
        const int b_vec = b_boxes * TN_BOX_BYTES / 16;
        for (int it = 0; it < steps; it++) {
            const int s = it % TN_STAGES;
            if (!mbar_wait(&bars->full[s], (it / TN_STAGES) & 1, err)) goto teardown;
            uint8_t *st = smem + (size_t)s * TN_STAGE_BYTES;
            auto split = [&](float4 *big, float4 *sm, int n_vec) {
                for (int i = tid; i < n_vec; i += 128) {
                    const float4 v = big[i];
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(v.x) & TF32_MASK); lo.x = __uint_as_float(__float_as_uint(v.x - hi.x) & TF32_MASK);
                    hi.y = __uint_as_float(__float_as_uint(v.y) & TF32_MASK); lo.y = __uint_as_float(__float_as_uint(v.y - hi.y) & TF32_MASK);
                    hi.z = __uint_as_float(__float_as_uint(v.z) & TF32_MASK); lo.z = __uint_as_float(__float_as_uint(v.z - hi.z) & TF32_MASK);
                    hi.w = __uint_as_float(__float_as_uint(v.w) & TF32_MASK); lo.w = __uint_as_float(__float_as_uint(v.w - hi.w) & TF32_MASK);
                    big[i] = hi;
                    sm[i] = lo;
                }
            };
            if (keep) {
                // A: four [32 rows x 128 bytes] boxes, 32-byte chunks of a row XORed with row%4 (SWIZZLE_128B_ATOM_32B)
                float4 *a4 = reinterpret_cast<float4 *>(st);
                for (int i = tid; i < TN_OPERAND_BYTES / 16; i += 128) {
                    const int fg = i >> 8, in_box = i & 255, row = in_box >> 3, sub = in_box & 7;
                    const int col = m_tile * 128 + fg * 32 + ((((sub >> 1) ^ (row & 3)) << 3) | ((sub & 1) << 2));
                    const int grow = r_lo + it * TN_ROWS + row;
                    uint32_t nib = 0;
                    if (grow < M && col < KA) {
                        const size_t bit = (size_t)grow * KA + col;
                        const uint32_t lo32 = keep[bit >> 5], hi32 = (bit & 31) > 28 ? keep[(bit >> 5) + 1] : 0u;
                        nib = __funnelshift_r(lo32, hi32, (uint32_t)(bit & 31)) & 0xfu;
                    }
                    float4 v = a4[i];
                    v.x = (nib & 1u) ? v.x : 0.f; v.y = (nib & 2u) ? v.y : 0.f; v.z = (nib & 4u) ? v.z : 0.f; v.w = (nib & 8u) ? v.w : 0.f;
                    a4[i] = v;
                }
                __syncwarp();
            }
            split(reinterpret_cast<float4 *>(st), reinterpret_cast<float4 *>(st + TN_OPERAND_BYTES), TN_OPERAND_BYTES / 16);
            split(reinterpret_cast<float4 *>(st + 2 * TN_OPERAND_BYTES), reinterpret_cast<float4 *>(st + 3 * TN_OPERAND_BYTES), b_vec);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->ready[s]);
        }
    } else if (warp >= 8) {
        // ======================= epilogue: the partial tile of this part -> workspace =======================
        const int q = warp - 8;
        if (!mbar_wait(&bars->acc_full, 0, err)) goto teardown;
        tc_fence_after();
        float *out = ws + ((size_t)part * rows_ws + m_tile * 128 + q * 32 + lane) * ld_ws + n_tile * nt_cols;
        for (int c0 = 0; c0 < nt_cols; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4 *>(out + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
        tc_fence_before();
    }
teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// C[r, c] (pitch ldc) = sum over parts of ws[part][r][c], in part order
__global__ void reduce_tn_kernel(const float *__restrict__ ws, float *__restrict__ c, int ka, int n, int ld_ws, int rows_ws, int ldc, int parts,
                                 float out_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ka * n) return;
    const int r = i / n, cc = i % n;
    float s0 = 0.f, s1 = 0.f;
    int b = 0;
    for (; b + 1 < parts; b += 2) {
        s0 += ws[((size_t)b * rows_ws + r) * ld_ws + cc];
        s1 += ws[((size_t)(b + 1) * rows_ws + r) * ld_ws + cc];
    }
    if (b < parts) s0 += ws[((size_t)b * rows_ws + r) * ld_ws + cc];
    c[(size_t)r * ldc + cc] = out_scale * (s0 + s1);
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

static bool make_tensor_map_2d_mode(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint64_t pitch_floats, uint32_t box_rows,
                                    uint32_t box_cols, CUtensorMapSwizzle swizzle) {
    EncodeTiled fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {pitch_floats * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool make_tensor_map_2d(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint64_t pitch_floats, uint32_t box_rows,
                        uint32_t box_cols, bool swizzle128) {
    return make_tensor_map_2d_mode(map, base, rows, cols, pitch_floats, box_rows, box_cols,
                                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

// ------------------------------------------------------------------------------------ nn / nt ----
static bool tc_disabled() {
    static const bool off = getenv("GCNK_NO_TCGEN05") && atoi(getenv("GCNK_NO_TCGEN05")) != 0;
    return off;
}
// n tile: all of n when its padding fits 64 columns, else 64-column tiles
static int nn_tile(int n) { const int npad = (n + 15) / 16 * 16; return npad <= 64 ? npad : 64; }
static size_t nn_smem(int k, int nt_cols) {
    const int kpad = (k + BK - 1) / BK * BK;
    return 2 * (size_t)STAGES * A_TILE_BYTES + 2 * (size_t)nt_cols * kpad * 4 + 4 * 32 * (size_t)epi_stride(nt_cols) * sizeof(float) + sizeof(Bars) + 1024;
}

// shapes the kernel takes: enough rows to matter, A's row pitch a multiple of 16 bytes, both split copies of the CTA's B tile resident
bool matmul_tc_nn_supported(int m, int k, int n, int lda, int ldc, bool b_is_nk) {
    (void)ldc; (void)b_is_nk;
    if (tc_disabled() || m < 1024 || k < 16 || lda % 4 || n < 8) return false;
    return nn_smem(k, nn_tile(n)) <= 227 * 1024 && tensor_maps_available();
}
bool matmul_tc_supported(int m, int k, int n) { return k >= 64 && k % 4 == 0 && n <= 64 && matmul_tc_nn_supported(m, k, n, k, n, false); }

int matmul_tc_nn(const float *a, int lda, const float *b, int ldb, bool b_is_nk, float *c, int ldc, int m, int k, int n, const float *row_scale,
                 cudaStream_t st, const uint32_t *keep, float out_scale, int relu) {
    const int nt_cols = nn_tile(n), ntn = (n + nt_cols - 1) / nt_cols, npad_all = ntn * nt_cols;
    const int kpad = (k + BK - 1) / BK * BK, kblocks = kpad / BK;
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    static float *bt[64] = {nullptr};
    static size_t bt_elems[64] = {0};
    const size_t need = 2 * (size_t)npad_all * kpad;
    if (bt_elems[dev] < need) {
        GCNK_CUDA(cudaStreamSynchronize(st));
        if (bt[dev]) GCNK_CUDA(cudaFree(bt[dev]));
        GCNK_CUDA(cudaMalloc(&bt[dev], sizeof(float) * need));
        bt_elems[dev] = need;
    }
    float *bt_big = bt[dev], *bt_small = bt[dev] + (size_t)npad_all * kpad;
    prep_b_kernel<<<(npad_all * kpad + 255) / 256, 256, 0, st>>>(b, ldb, b_is_nk ? 1 : 0, bt_big, bt_small, k, n, kpad, npad_all);
    GCNK_LAUNCHED();
    CUtensorMap map_a, map_bb, map_bs;
    if (!make_tensor_map_2d(&map_a, a, (uint64_t)m, (uint64_t)k, (uint64_t)lda, BM, BK, true) ||
        !make_tensor_map_2d(&map_bb, bt_big, (uint64_t)npad_all, (uint64_t)kpad, (uint64_t)kpad, (uint32_t)nt_cols, BK, true) ||
        !make_tensor_map_2d(&map_bs, bt_small, (uint64_t)npad_all, (uint64_t)kpad, (uint64_t)kpad, (uint32_t)nt_cols, BK, true)) {
        set_error("matmul_tc: cuTensorMapEncodeTiled failed");
        return GCNK_EUNSUPPORTED;
    }
    const size_t smem = nn_smem(k, nt_cols);
    static bool attr[64] = {false};
    if (!attr[dev]) {
        GCNK_CUDA(cudaFuncSetAttribute(matmul_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr[dev] = true;
    }
    const int n_tiles = (m + BM - 1) / BM;
    const int grid = std::max(1, std::min(n_tiles, sm_count() / ntn)) * ntn;      // a multiple of ntn: every CTA owns one n tile
    matmul_tc_kernel<<<grid, TC_THREADS, smem, st>>>(map_a, map_bb, map_bs, c, ldc, row_scale, m, n, nt_cols, ntn, kblocks, keep, k, out_scale, relu,
                                                     async_err_flag());
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int matmul_tc_fw(const float *a, const float *b, float *c, int m, int k, int n, cudaStream_t st) {
    return matmul_tc_nn(a, k, b, n, false, c, n, m, k, n, nullptr, st, nullptr, 1.0f, 0);
}

// ----------------------------------------------------------------------------------------- tn ----
static int tn_tile_n(int n) { const int npad = (n + 15) / 16 * 16; return npad <= 128 ? npad : 128; }
static int tn_parts(int m, int tiles) { return std::max(1, std::min(sm_count() / tiles, (m + 4 * TN_ROWS - 1) / (4 * TN_ROWS))); }
bool matmul_tc_tn_supported(int m, int ka, int n, int lda, int ldb) {
    if (tc_disabled() || m < 2048 || lda % 4 || ldb % 4 || ka < 8 || n < 8) return false;
    const int tiles = ((ka + 127) / 128) * ((n + tn_tile_n(n) - 1) / tn_tile_n(n));
    return tiles <= sm_count() && tensor_maps_available();
}
size_t matmul_tc_tn_workspace(int m, int ka, int n) {
    const int nt_cols = tn_tile_n(n), mt = (ka + 127) / 128, ntn = (n + nt_cols - 1) / nt_cols;
    return sizeof(float) * (size_t)tn_parts(m, mt * ntn) * (mt * 128) * (ntn * nt_cols);
}

int matmul_tc_tn(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int ka, int n, float *ws, size_t ws_bytes, cudaStream_t st,
                 const uint32_t *keep, float out_scale) {
    const int nt_cols = tn_tile_n(n), mt = (ka + 127) / 128, ntn = (n + nt_cols - 1) / nt_cols, tiles = mt * ntn;
    const int parts = tn_parts(m, tiles);
    if (!ws || ws_bytes < matmul_tc_tn_workspace(m, ka, n)) { set_error("matmul_tc_tn: workspace too small"); return GCNK_EINVAL; }
    CUtensorMap map_a, map_b;
    if (!make_tensor_map_2d_mode(&map_a, a, (uint64_t)m, (uint64_t)ka, (uint64_t)lda, TN_ROWS, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_tensor_map_2d_mode(&map_b, b, (uint64_t)m, (uint64_t)n, (uint64_t)ldb, TN_ROWS, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
        set_error("matmul_tc_tn: cuTensorMapEncodeTiled failed");
        return GCNK_EUNSUPPORTED;
    }
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    static bool attr[64] = {false};
    if (!attr[dev]) {
        GCNK_CUDA(cudaFuncSetAttribute(matmul_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr[dev] = true;
    }
    const int ld_ws = ntn * nt_cols, rows_ws = mt * 128;
    // rows of m per part: whole 32-row stages
    int rows_per_part = ((m + parts - 1) / parts + TN_ROWS - 1) / TN_ROWS * TN_ROWS;
    const int parts_used = (m + rows_per_part - 1) / rows_per_part;
    const size_t smem = (size_t)TN_STAGES * TN_STAGE_BYTES + sizeof(TnBars) + 1024;
    matmul_tn_kernel<<<tiles * parts_used, TC_THREADS, smem, st>>>(map_a, map_b, ws, ld_ws, rows_ws, m, rows_per_part, mt, ntn, nt_cols, keep, ka,
                                                                   async_err_flag());
    GCNK_LAUNCHED();
    reduce_tn_kernel<<<(ka * n + 255) / 256, 256, 0, st>>>(ws, c, ka, n, ld_ws, rows_ws, ldc, parts_used, out_scale);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // namespace gcnk
