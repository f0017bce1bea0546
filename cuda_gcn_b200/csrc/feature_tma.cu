// feature_tma.cu — the dense 602 -> 16 feature transform and its weight gradient with the X stream staged by TMA.
//
// Same math and the same 3xTF32 mma.sync fragments as feature_tc.cu, but the operand stream no longer lives in
// registers: a producer lane issues cp.async.bulk.tensor loads of X boxes into a deep shared-memory ring
// (128 KB forward / 195 KB backward in flight per SM) and the consumer warps read their fragments from there.
// feature_tc.cu holds at most one k-step per warp in flight (registers) and stalls 50 % of the time on the X loads.
//
// TMA needs a row pitch that is a multiple of 16 bytes; the reference's layout (602 floats = 2,408 B) is not, so
// these kernels take a PACKED copy of the matrix with a padded pitch (gcnk_dense_pack, made once: the feature matrix
// is constant over the whole training run).  Boxes are 32 floats wide with the 128-byte swizzle, so a fragment
// read (lane (g,t): rows g / g+8, columns 8j+2t, +1) touches 16 distinct 8-byte slots per half-warp.
// The keep bits stay in the reference's flat order (bit row*n + col of the UNPADDED matrix).
#include <stdlib.h>

#include <algorithm>

#include "tma.cuh"

using namespace gcnk;

namespace {

constexpr int P = 16;
constexpr uint32_t TF32_MASK = 0xffffe000u;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_rn(float x, uint32_t &big, uint32_t &small) {
    big = to_tf32(x);
    small = to_tf32(x - __uint_as_float(big));
}
__device__ __forceinline__ void split_trunc(float x, uint32_t &big, uint32_t &small) {
    big = __float_as_uint(x) & TF32_MASK;
    small = __float_as_uint(x - __uint_as_float(big)) & TF32_MASK;
}
__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// Keep-bit words travel global -> shared memory with 4-byte cp.async, several ring slots ahead of their use: a register
// prefetch one step ahead did not cover the L2 round trip (measured with the loads removed: 23 us of the forward and 25 us
// of the backward kernel were exposed keep-bit latency).  One group per step; a slot is read after wait_group + __syncwarp.
constexpr int KW_SLOTS = 4, KW_AHEAD = 3;
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// ================================================================================== forward ====
// CTA = 16 consumer warps (16 rows each: a 256-row tile) + 1 producer warp.  Stage = one [256 x 32] box (32 KB).
// (8 consumer warps per SM were measured slower than the register-staged kernel: too few warps to hide the
// shared-memory and MMA latencies of the dependent k-step chains.)
// The tile height is a template parameter (11 .. 16 consumer warps), chosen by fw_pick_consumers.
constexpr int FW_MAX_CONSUMERS = 16, FW_MIN_CONSUMERS = 11, FW_STAGES = 4;

struct FwBars { uint64_t full[FW_STAGES], empty[FW_STAGES]; };

template <int FW_CONSUMERS>
__global__ void __launch_bounds__((FW_CONSUMERS + 1) * 32, 1) dense_fw16_tma_kernel(const __grid_constant__ CUtensorMap map_x, const float *__restrict__ w,
                                                                        float *__restrict__ c, int m, int n,
                                                                        const uint32_t *__restrict__ bits, int64_t bit_words, float scale,
                                                                        const float *__restrict__ row_scale, int relu, int *err) {
    constexpr int FW_THREADS = (FW_CONSUMERS + 1) * 32, FW_BM = 16 * FW_CONSUMERS, FW_STAGE_BYTES = FW_BM * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer (LDS, not LD)
    const int KS = (n + 7) / 8, n_batches = (n + 31) / 32;
    uint8_t *ring = smem;                                                        // [FW_STAGES][FW_BM x 128 B]
    float4 *sfrag = reinterpret_cast<float4 *>(smem + FW_STAGES * FW_STAGE_BYTES);   // [KS][32] big, [KS][32] small
    FwBars *bars = reinterpret_cast<FwBars *>(sfrag + 2 * KS * 32);
    uint32_t *kws_all = reinterpret_cast<uint32_t *>(bars + 1);                  // [FW_CONSUMERS][KW_SLOTS][16 rows][2 words]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, g = lane >> 2;
    for (int i = threadIdx.x; i < KS * 32; i += FW_THREADS) {
        const int s = i >> 5, l = i & 31, tt = l & 3, gg = l >> 2;
        const int k0 = 8 * s + 2 * tt, k1 = k0 + 1;
        const float v[4] = {k0 < n ? w[k0 * P + gg] : 0.f, k1 < n ? w[k1 * P + gg] : 0.f,
                            k0 < n ? w[k0 * P + 8 + gg] : 0.f, k1 < n ? w[k1 * P + 8 + gg] : 0.f};
        uint32_t big[4], small[4];
#pragma unroll
        for (int e = 0; e < 4; e++) split_rn(v[e], big[e], small[e]);
        sfrag[i] = make_float4(__uint_as_float(big[0]), __uint_as_float(big[1]), __uint_as_float(big[2]), __uint_as_float(big[3]));
        sfrag[KS * 32 + i] = make_float4(__uint_as_float(small[0]), __uint_as_float(small[1]), __uint_as_float(small[2]), __uint_as_float(small[3]));
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < FW_STAGES; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], FW_CONSUMERS); }
        mbar_fence_init();
    }
    __syncthreads();

    const int n_tiles = (m + FW_BM - 1) / FW_BM;
    if (warp == FW_CONSUMERS) {
        // ------------------------------------------------ producer ------------------------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int b = 0; b < n_batches; b++, it++) {
                    const int s = it % FW_STAGES;
                    if (it >= FW_STAGES && !mbar_wait(&bars->empty[s], ((it / FW_STAGES) - 1) & 1, err)) return;
                    mbar_expect_tx(&bars->full[s], FW_STAGE_BYTES);
                    tma_load_2d(ring + s * FW_STAGE_BYTES, &map_x, &bars->full[s], 32 * b, tile * FW_BM);
                }
        }
        return;
    }
    // -------------------------------------------------- consumers --------------------------------------------------
    uint32_t it = 0;
    // keep words: lane l fetches word (l & 1) of the window of tile-local row 16 * warp + (l >> 1) (rows g and g + 8 of the
    // lanes' fragments), KW_AHEAD batches ahead of the batch being multiplied, across tile boundaries
    uint32_t *kws = kws_all + warp * (KW_SLOTS * 32);
    int p_tile = blockIdx.x, p_b = 0;
    uint32_t p_it = 0;
    auto prefetch_keep = [&]() {
        if (bits && p_tile < n_tiles) {
            const int row = min(p_tile * FW_BM + 16 * warp + (lane >> 1), m - 1);
            const int64_t w = (((int64_t)row * n + 32 * p_b) >> 5) + (lane & 1);
            cp_async_4(kws + (p_it % KW_SLOTS) * 32 + lane, bits + (w < bit_words ? w : bit_words - 1));
            if (++p_b == n_batches) { p_b = 0; p_tile += gridDim.x; }
        }
        p_it++;
        cp_async_commit();
    };
    if (bits)
        for (int k = 0; k < KW_AHEAD; k++) prefetch_keep();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int lr = 16 * warp + g;                                  // tile-local rows lr and lr + 8 (both have (row & 7) == g)
        const int ra = tile * FW_BM + lr, rb = ra + 8;
        const bool va = ra < m, vb = rb < m;
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
        // bit offset of the rows' windows inside their first word (the same for every batch: batches advance by 32 bits)
        const uint32_t sha = (uint32_t)(((int64_t)min(ra, m - 1) * n) & 31), shb = (uint32_t)(((int64_t)min(rb, m - 1) * n) & 31);
#pragma unroll 1
        for (int b = 0; b < n_batches; b++, it++) {
            const int s = it % FW_STAGES;
            uint32_t wa = 0, wb = 0;
            if (bits) {
                prefetch_keep();
                cp_async_wait<KW_AHEAD>();                             // this batch's group has landed (for the lane that issued it)
                __syncwarp();
                const uint32_t *slot = kws + (it % KW_SLOTS) * 32;
                const uint2 qa = *reinterpret_cast<const uint2 *>(slot + 2 * g), qb = *reinterpret_cast<const uint2 *>(slot + 2 * (8 + g));
                wa = va ? __funnelshift_r(qa.x, qa.y, sha) : 0u;
                wb = vb ? __funnelshift_r(qb.x, qb.y, shb) : 0u;
            }
            if (!mbar_wait(&bars->full[s], (it / FW_STAGES) & 1, err)) return;
            const uint8_t *st = ring + s * FW_STAGE_BYTES;
            const uint32_t ka = wa >> (2 * t), kb = wb >> (2 * t);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int ks = 4 * b + j;
                if (ks < KS) {
                    // columns 8j+2t, +1: 16-byte chunk 2j + (t >> 1), XOR-swizzled with the row's low 3 bits (= g)
                    const uint32_t off = (uint32_t)((((2 * j + (t >> 1)) ^ g) << 4) | ((t & 1) << 3));
                    float2 fa = *reinterpret_cast<const float2 *>(st + lr * 128 + off);
                    float2 fb = *reinterpret_cast<const float2 *>(st + (lr + 8) * 128 + off);
                    if (bits) {                                         // 1/(1-p) is applied once, in the epilogue
                        fa.x = ka & (1u << (8 * j)) ? fa.x : 0.f;
                        fa.y = ka & (2u << (8 * j)) ? fa.y : 0.f;
                        fb.x = kb & (1u << (8 * j)) ? fb.x : 0.f;
                        fb.y = kb & (2u << (8 * j)) ? fb.y : 0.f;
                    }
                    uint32_t ab[4], as[4];
                    split_trunc(fa.x, ab[0], as[0]);
                    split_trunc(fb.x, ab[1], as[1]);
                    split_trunc(fa.y, ab[2], as[2]);
                    split_trunc(fb.y, ab[3], as[3]);
                    const float4 wb4 = sfrag[ks * 32 + lane], ws4 = sfrag[(KS + ks) * 32 + lane];
                    mma(acc0, as[0], as[1], as[2], as[3], __float_as_uint(wb4.x), __float_as_uint(wb4.y));
                    mma(acc1, as[0], as[1], as[2], as[3], __float_as_uint(wb4.z), __float_as_uint(wb4.w));
                    mma(acc0, ab[0], ab[1], ab[2], ab[3], __float_as_uint(ws4.x), __float_as_uint(ws4.y));
                    mma(acc1, ab[0], ab[1], ab[2], ab[3], __float_as_uint(ws4.z), __float_as_uint(ws4.w));
                    mma(acc0, ab[0], ab[1], ab[2], ab[3], __float_as_uint(wb4.x), __float_as_uint(wb4.y));
                    mma(acc1, ab[0], ab[1], ab[2], ab[3], __float_as_uint(wb4.z), __float_as_uint(wb4.w));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->empty[s]);               // this warp is done with the slot
        }
        if (relu) {
#pragma unroll
            for (int e = 0; e < 4; e++) { acc0[e] = acc0[e] > 0.f ? acc0[e] : 0.f; acc1[e] = acc1[e] > 0.f ? acc1[e] : 0.f; }
        }
        const float post = bits ? scale : 1.f;
        if (va) {
            const float rs = post * (row_scale ? row_scale[ra] : 1.f);
            float *o = c + (size_t)ra * P + 2 * t;
            *reinterpret_cast<float2 *>(o) = make_float2(rs * acc0[0], rs * acc0[1]);
            *reinterpret_cast<float2 *>(o + 8) = make_float2(rs * acc1[0], rs * acc1[1]);
        }
        if (vb) {
            const float rs = post * (row_scale ? row_scale[rb] : 1.f);
            float *o = c + (size_t)rb * P + 2 * t;
            *reinterpret_cast<float2 *>(o) = make_float2(rs * acc0[2], rs * acc0[3]);
            *reinterpret_cast<float2 *>(o + 8) = make_float2(rs * acc1[2], rs * acc1[3]);
        }
    }
}

// ================================================================================= backward ====
// dW^T[16 x n] = G^T[16 x m] * drop(X)[m x n].  A CTA owns a slab of rows; consumer warp w owns the 64 features of
// boxes 2w and 2w+1 (4 pairs of n-tiles).  Stage = 16 rows = two k-steps: the row block as n_boxes [16 x 32] boxes
// (2 KB each, 128-byte swizzle) plus the [16 x 16] block of G (1 KB, no swizzle).
constexpr int BW_MAX_BOXES = 20, BW_ROWS = 16, BW_BOX_BYTES = BW_ROWS * 128, BW_G_BYTES = BW_ROWS * P * 4, BW_STAGES = 5;

struct BwBars { uint64_t full[BW_STAGES], empty[BW_STAGES]; };

constexpr int BW_THREADS = 21 * 32;     // two groups of 10 consumer warps (k-step 0 / k-step 1 of every stage) + the producer warp

__global__ void __launch_bounds__(BW_THREADS, 1) dense_bw16_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                                                                 float *__restrict__ partials, int m, int n, int n_boxes, int rows_per_cta,
                                                                 const uint32_t *__restrict__ bits, int64_t bit_words, float scale, int *err) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer (LDS, not LD)
    const uint32_t stage_bytes = (uint32_t)n_boxes * BW_BOX_BYTES + 1024;          // X boxes, then G (padded to keep 1 KB alignment)
    BwBars *bars = reinterpret_cast<BwBars *>(smem + BW_STAGES * stage_bytes);
    uint32_t *kws_all = reinterpret_cast<uint32_t *>(bars + 1);                    // [20 warps][KW_SLOTS][8 rows][4 words]
    const int n_consumers = (n_boxes + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3, g = lane >> 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_STAGES; s++) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 2 * n_consumers); }
        mbar_fence_init();
    }
    __syncthreads();
    const int r_lo = blockIdx.x * rows_per_cta, r_hi = min(m, r_lo + rows_per_cta);
    const int n_stages = r_lo < r_hi ? (r_hi - r_lo + BW_ROWS - 1) / BW_ROWS : 0;

    if (warp == 20) {
        if (lane == 0) {
            for (int it = 0; it < n_stages; it++) {
                const int s = it % BW_STAGES;
                if (it >= BW_STAGES && !mbar_wait(&bars->empty[s], ((it / BW_STAGES) - 1) & 1, err)) return;
                uint8_t *st = smem + s * stage_bytes;
                mbar_expect_tx(&bars->full[s], (uint32_t)n_boxes * BW_BOX_BYTES + BW_G_BYTES);
                const int row = r_lo + it * BW_ROWS;
                for (int bx = 0; bx < n_boxes; bx++) tma_load_2d(st + bx * BW_BOX_BYTES, &map_x, &bars->full[s], 32 * bx, row);
                tma_load_2d(st + n_boxes * BW_BOX_BYTES, &map_g, &bars->full[s], 0, row);
            }
        }
        return;
    }
    const int group = warp / 10, bw = warp % 10;          // group = which k-step of a stage; bw = feature band
    if (bw >= n_consumers) return;

    const int f_band = 64 * bw;
    float acc[4][2][4];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[p][h][e] = 0.f;

    // keep words of the stage's 8 rows of this warp's group, 64 + 31 bits from the band's first feature on: lane l fetches word
    // (l & 3) of row (l >> 2), KW_AHEAD stages ahead
    uint32_t *kws = kws_all + warp * (KW_SLOTS * 32);
    int p_it = 0;
    auto prefetch_keep = [&]() {
        if (bits && p_it < n_stages) {
            const int row = min(r_lo + p_it * BW_ROWS + 8 * group + (lane >> 2), m - 1);
            const int64_t w = (((int64_t)row * n + f_band) >> 5) + (lane & 3);
            cp_async_4(kws + (p_it % KW_SLOTS) * 32 + lane, bits + (w < bit_words ? w : bit_words - 1));
        }
        p_it++;
        cp_async_commit();
    };
    if (bits)
        for (int k = 0; k < KW_AHEAD; k++) prefetch_keep();
#pragma unroll 1
    for (int it = 0; it < n_stages; it++) {
        const int s = it % BW_STAGES;
        // keep windows of this lane's rows for both k-steps of the stage: 64 + 31 bits -> two windows of 32 out of three words
        uint32_t kw[2][2];
        if (bits) {
            prefetch_keep();
            cp_async_wait<KW_AHEAD>();
            __syncwarp();
            const uint32_t *slot = kws + (it % KW_SLOTS) * 32;
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const int row = r_lo + it * BW_ROWS + 8 * group + 2 * t + rr;
                const uint4 q = *reinterpret_cast<const uint4 *>(slot + 4 * (2 * t + rr));
                const uint32_t sh = (uint32_t)(((int64_t)min(row, m - 1) * n + f_band) & 31);
                kw[rr][0] = row < r_hi ? __funnelshift_r(q.x, q.y, sh) : 0u;
                kw[rr][1] = row < r_hi ? __funnelshift_r(q.y, q.z, sh) : 0u;
            }
        }
        if (!mbar_wait(&bars->full[s], (it / BW_STAGES) & 1, err)) return;
        const uint8_t *st = smem + s * stage_bytes;
        const float *gs = reinterpret_cast<const float *>(st + n_boxes * BW_BOX_BYTES);          // [16 rows][16]
        {
            const int lr0 = 8 * group + 2 * t, lr1 = lr0 + 1;                                    // stage-local rows of this lane
            uint32_t ab[4], as[4];
            split_trunc(gs[lr0 * P + g], ab[0], as[0]);
            split_trunc(gs[lr0 * P + g + 8], ab[1], as[1]);
            split_trunc(gs[lr1 * P + g], ab[2], as[2]);
            split_trunc(gs[lr1 * P + g + 8], ab[3], as[3]);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int box = 2 * bw + (p >> 1), cc = 16 * (p & 1) + 2 * g;                    // column inside the 32-wide box
                if (box < n_boxes) {
                    const uint8_t *bx = st + box * BW_BOX_BYTES;
                    float2 f0 = *reinterpret_cast<const float2 *>(bx + sw128_offset(lr0, cc));
                    float2 f1 = *reinterpret_cast<const float2 *>(bx + sw128_offset(lr1, cc));
                    if (bits) {                                                                  // 1/(1-p) goes onto the partial sums
                        const int sh = 16 * (p & 1) + 2 * g;                                     // bit inside the 32-bit window of this box
                        const uint32_t w0 = kw[0][p >> 1], w1 = kw[1][p >> 1];
                        f0.x = (w0 >> sh) & 1u ? f0.x : 0.f;
                        f0.y = (w0 >> sh) & 2u ? f0.y : 0.f;
                        f1.x = (w1 >> sh) & 1u ? f1.x : 0.f;
                        f1.y = (w1 >> sh) & 2u ? f1.y : 0.f;
                    }
                    uint32_t bb[4], bs[4];
                    split_trunc(f0.x, bb[0], bs[0]);
                    split_trunc(f1.x, bb[1], bs[1]);
                    split_trunc(f0.y, bb[2], bs[2]);
                    split_trunc(f1.y, bb[3], bs[3]);
                    mma(acc[p][0], as[0], as[1], as[2], as[3], bb[0], bb[1]);
                    mma(acc[p][1], as[0], as[1], as[2], as[3], bb[2], bb[3]);
                    mma(acc[p][0], ab[0], ab[1], ab[2], ab[3], bs[0], bs[1]);
                    mma(acc[p][1], ab[0], ab[1], ab[2], ab[3], bs[2], bs[3]);
                    mma(acc[p][0], ab[0], ab[1], ab[2], ab[3], bb[0], bb[1]);
                    mma(acc[p][1], ab[0], ab[1], ab[2], ab[3], bb[2], bb[3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->empty[s]);
    }
    // D fragment of an n-tile: c0 = (h g, n 2t), c1 = (h g, n 2t+1), c2 = (h g+8, n 2t), c3 = (h g+8, n 2t+1); logical column
    // n of the even tile of a pair is feature f0 + 2n, of the odd tile f0 + 2n + 1
    const float post = bits ? scale : 1.f;
    float *out = partials + ((size_t)blockIdx.x * 2 + group) * n * P;
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int fa = f_band + 16 * p + 2 * (2 * t) + h, fb = f_band + 16 * p + 2 * (2 * t + 1) + h;
            if (fa < n) { out[(size_t)fa * P + g] = post * acc[p][h][0]; out[(size_t)fa * P + g + 8] = post * acc[p][h][2]; }
            if (fb < n) { out[(size_t)fb * P + g] = post * acc[p][h][1]; out[(size_t)fb * P + g + 8] = post * acc[p][h][3]; }
        }
}

__global__ void __launch_bounds__(256) reduce_parts_tma_kernel(const float *__restrict__ partials, float *__restrict__ out, int elems, int parts) {
    reduce_parts_block(partials, out, elems, parts);
}

__global__ void pack_rows_kernel(const float *__restrict__ x, float *__restrict__ xp, int64_t m, int n, int ld) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = m * ld, stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int64_t r = i / ld;
        const int c = (int)(i - r * ld);
        xp[i] = c < n ? x[r * n + c] : 0.f;
    }
}

}  // namespace

// Consumer warps (tile height / 16) of the forward kernel for m rows.  Measured at 232,965 rows (r03e): 16 warps 0.282 ms per
// step, 13 warps 0.309, 11 warps 0.300 although 11 turns 6.15 rounds-run-as-7 into 8.95-run-as-9 — the warps hide more
// latency than the last round costs, so several rounds always run at 16.  A matrix that fits ONE round (a rank's slice of a
// row partition) takes the smallest tile that still fits one round: same number of rounds, fewer rows per SM.
// GCN_FW_WARPS=11..16 overrides.
int fw_pick_consumers(int m) {
    if (const char *e = getenv("GCN_FW_WARPS")) { const int v = atoi(e); if (v >= FW_MIN_CONSUMERS && v <= FW_MAX_CONSUMERS) return v; }
    const int sms = sm_count();
    if ((m + 16 * FW_MAX_CONSUMERS - 1) / (16 * FW_MAX_CONSUMERS) > sms) return FW_MAX_CONSUMERS;
    for (int cw = FW_MIN_CONSUMERS; cw < FW_MAX_CONSUMERS; cw++)
        if ((m + 16 * cw - 1) / (16 * cw) <= sms) return cw;
    return FW_MAX_CONSUMERS;
}

template <int CW>
int fw_launch(int grid, size_t smem, cudaStream_t st, const CUtensorMap &map_x, const float *w, float *c, int m, int n, const uint32_t *bits,
              int64_t words, float scale, const float *row_scale, int relu) {
    static bool attr[64] = {false};
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    if (!attr[dev]) {
        GCNK_CUDA(cudaFuncSetAttribute(dense_fw16_tma_kernel<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr[dev] = true;
    }
    dense_fw16_tma_kernel<CW><<<grid, (CW + 1) * 32, smem, st>>>(map_x, w, c, m, n, bits, words, scale, row_scale, relu, async_err_flag());
    return GCNK_OK;
}

extern "C" {

int gcnk_dense_pack(const float *x, int m, int n, float *xp, int ld, gcnk_stream_t stream) {
    GCNK_REQUIRE(x && xp && m >= 0 && n > 0 && ld >= n, "bad arguments");
    if (m == 0) return GCNK_OK;
    const int64_t total = (int64_t)m * ld;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    pack_rows_kernel<<<grid, 256, 0, S(stream)>>>(x, xp, m, n, ld);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_dense_transform_ld(const float *xp, int ld, int m, int n, const float *w, float *c, int p, const uint32_t *drop_bits,
                            float drop_scale, const float *row_scale, int relu, gcnk_stream_t stream) {
    GCNK_REQUIRE(xp && w && c && m >= 0 && n > 0 && ld >= n && p > 0, "bad arguments");
    if (m == 0) return GCNK_OK;
    const int KS = (n + 7) / 8;
    const int consumers = fw_pick_consumers(m);
    const int bm = 16 * consumers;
    const size_t smem = (size_t)FW_STAGES * bm * 128 + sizeof(float4) * 2 * (size_t)KS * 32 + sizeof(FwBars) + 1024 +
                        sizeof(uint32_t) * consumers * KW_SLOTS * 32;
    if (p != P || ld % 4 || reinterpret_cast<uintptr_t>(xp) % 16 || reinterpret_cast<uintptr_t>(c) % 8 || smem > 227 * 1024 ||
        !tensor_maps_available()) {
        set_error("gcnk_dense_transform_ld: needs p == 16, a 16-byte aligned pitch and base, n <= ~760 (got n=%d ld=%d p=%d)", n, ld, p);
        return GCNK_EUNSUPPORTED;
    }
    CUtensorMap map_x;
    if (!make_tensor_map_2d(&map_x, xp, (uint64_t)m, (uint64_t)n, (uint64_t)ld, bm, 32, true)) {
        set_error("gcnk_dense_transform_ld: cuTensorMapEncodeTiled failed");
        return GCNK_EUNSUPPORTED;
    }
    const int n_tiles = (m + bm - 1) / bm;
    const int grid = std::min(n_tiles, sm_count());
    const int64_t words = ((int64_t)m * n + 31) / 32;
    int rc = GCNK_OK;
    switch (consumers) {
        case 11: rc = fw_launch<11>(grid, smem, S(stream), map_x, w, c, m, n, drop_bits, words, drop_scale, row_scale, relu); break;
        case 12: rc = fw_launch<12>(grid, smem, S(stream), map_x, w, c, m, n, drop_bits, words, drop_scale, row_scale, relu); break;
        case 13: rc = fw_launch<13>(grid, smem, S(stream), map_x, w, c, m, n, drop_bits, words, drop_scale, row_scale, relu); break;
        case 14: rc = fw_launch<14>(grid, smem, S(stream), map_x, w, c, m, n, drop_bits, words, drop_scale, row_scale, relu); break;
        case 15: rc = fw_launch<15>(grid, smem, S(stream), map_x, w, c, m, n, drop_bits, words, drop_scale, row_scale, relu); break;
        default: rc = fw_launch<16>(grid, smem, S(stream), map_x, w, c, m, n, drop_bits, words, drop_scale, row_scale, relu); break;
    }
    if (rc) return rc;
    GCNK_LAUNCHED();
    return GCNK_OK;
}

size_t gcnk_dense_transform_bw_workspace(int m, int n) {
    const int ctas = std::max(1, std::min(sm_count(), (m + 15) / 16));
    return sizeof(float) * 2 * (size_t)ctas * n * P;       // two row groups per CTA
}

int gcnk_dense_transform_bw_ld(const float *xp, int ld, int m, int n, const float *g, float *w_grad, int p, const uint32_t *drop_bits,
                               float drop_scale, float *workspace, size_t workspace_bytes, gcnk_stream_t stream) {
    GCNK_REQUIRE(xp && g && w_grad && m > 0 && n > 0 && ld >= n && p > 0, "bad arguments");
    const int n_boxes = (n + 31) / 32;
    const size_t smem = (size_t)BW_STAGES * ((size_t)n_boxes * BW_BOX_BYTES + 1024) + sizeof(BwBars) + 1024 +
                        sizeof(uint32_t) * 20 * KW_SLOTS * 32;
    if (p != P || ld % 4 || reinterpret_cast<uintptr_t>(xp) % 16 || reinterpret_cast<uintptr_t>(g) % 16 || n_boxes > BW_MAX_BOXES ||
        smem > 227 * 1024 || !tensor_maps_available()) {
        set_error("gcnk_dense_transform_bw_ld: needs p == 16, a 16-byte aligned pitch and bases, n <= 640 (got n=%d ld=%d p=%d)", n, ld, p);
        return GCNK_EUNSUPPORTED;
    }
    GCNK_REQUIRE(workspace && workspace_bytes >= gcnk_dense_transform_bw_workspace(m, n), "workspace too small");
    CUtensorMap map_x, map_g;
    if (!make_tensor_map_2d(&map_x, xp, (uint64_t)m, (uint64_t)n, (uint64_t)ld, BW_ROWS, 32, true) ||
        !make_tensor_map_2d(&map_g, g, (uint64_t)m, (uint64_t)P, (uint64_t)P, BW_ROWS, P, false)) {
        set_error("gcnk_dense_transform_bw_ld: cuTensorMapEncodeTiled failed");
        return GCNK_EUNSUPPORTED;
    }
    static bool attr[64] = {false};
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    if (!attr[dev]) {
        GCNK_CUDA(cudaFuncSetAttribute(dense_bw16_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr[dev] = true;
    }
    int ctas = std::max(1, std::min(sm_count(), (m + 15) / 16));
    int rows_per_cta = ((m + ctas - 1) / ctas + BW_ROWS - 1) / BW_ROWS * BW_ROWS;
    ctas = (m + rows_per_cta - 1) / rows_per_cta;
    dense_bw16_tma_kernel<<<ctas, BW_THREADS, smem, S(stream)>>>(map_x, map_g, workspace, m, n, n_boxes, rows_per_cta, drop_bits,
                                                          ((int64_t)m * n + 31) / 32, drop_scale, async_err_flag());
    GCNK_LAUNCHED();
    const int elems = n * P;
    reduce_parts_tma_kernel<<<(elems + 31) / 32, 256, 0, S(stream)>>>(workspace, w_grad, elems, 2 * ctas);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // extern "C"
