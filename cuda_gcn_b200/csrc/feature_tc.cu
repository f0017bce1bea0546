// feature_tc.cu — the dense-feature SparseMatmul at hidden width 16 on the tensor cores, fp32-accurate.
//
//   forward   C[m x 16]  = drop(X)[m x n] * W[n x 16] (* row_scale)     (SparseMatmul::forward, module.cpp:47-61)
//   backward  dW[n x 16] = drop(X)^T * G[m x 16]                         (SparseMatmul::backward, module.cpp:63-77)
//
// for a feature matrix stored with every column present (Reddit as reddit_preprocess.py:161-167 dumps
// it: 602 entries per row).  Both are one streaming pass over X (561 MB at Reddit shape), i.e. HBM-bound
// at ~86 us; an fp32 SIMT formulation needs ~16 FMA + shared-memory operand traffic per loaded element
// and lands at 10x that (measured: 1.0 / 1.2 ms).  Here the contraction runs on mma.sync.m16n8k8 TF32
// with the 3xTF32 split (x = big + small, both TF32; big*big + big*small + small*big), which keeps ~22
// mantissa bits per operand — the products are as accurate as fp32 FMAs, accumulation is fp32.
//
// Why mma.sync and not tcgen05 here: the X rows are 602 floats = 2,408 bytes apart, not a multiple of
// 16 bytes, so neither a TMA tensor map nor cp.async.bulk can address them without a re-packed copy of
// X; UMMA would also need both split halves of every X tile materialised in shared memory in the
// canonical swizzled layout.  With register operands the split is three ALU ops per element and the
// operand fragments are loaded straight from global memory as 8-byte vectors.  The kernel is bound by
// the X stream, not by the tensor pipe.
//
// Fragment trick (both kernels): an MMA does not care which physical index a logical k (or n) index
// denotes as long as A and B agree, so logical k = t and k = t+4 are mapped to the ADJACENT physical
// columns 2t and 2t+1.  One 64-bit load then feeds two fragment registers, and a quad of lanes reads
// 32 contiguous bytes.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

using namespace gcnk;

namespace gcnk_tc {

constexpr int P = 16;                       // output width handled here
constexpr int THREADS = 256, WARPS = THREADS / 32;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split(float x, uint32_t &big, uint32_t &small) {
    big = to_tf32(x);
    small = to_tf32(x - __uint_as_float(big));      // exact difference, then rounded to TF32
}
// The same split for the streamed operand (140 M elements per pass), by truncation: two LOP3 and one FADD on the
// full-rate ALU pipe instead of two conversions.  big = x with the 13 low mantissa bits cleared (a valid TF32),
// small = x - big exactly, truncated again; what is lost is below 2^-20 of x.
constexpr uint32_t TF32_MASK = 0xffffe000u;
__device__ __forceinline__ void split_trunc(float x, uint32_t &big, uint32_t &small) {
    big = __float_as_uint(x) & TF32_MASK;
    small = __float_as_uint(x - __uint_as_float(big)) & TF32_MASK;
}
__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 3xTF32 with one accumulator per split term: the three MMAs of a k-step are independent, so the tensor
// pipe is not serialised on one accumulation chain; the terms are added once, in the epilogue
// (cross terms first, then the dominant one).
struct Acc3 {
    float bb[4], bs[4], sb[4];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int e = 0; e < 4; e++) bb[e] = bs[e] = sb[e] = 0.f;
    }
    __device__ __forceinline__ float total(int e) const { return (sb[e] + bs[e]) + bb[e]; }
};
__device__ __forceinline__ void mma3(Acc3 &d, const uint32_t (&ab)[4], const uint32_t (&as)[4], uint32_t bb0, uint32_t bb1,
                                     uint32_t bs0, uint32_t bs1) {
    mma(d.sb, as[0], as[1], as[2], as[3], bb0, bb1);
    mma(d.bs, ab[0], ab[1], ab[2], ab[3], bs0, bs1);
    mma(d.bb, ab[0], ab[1], ab[2], ab[3], bb0, bb1);
}
__device__ __forceinline__ float2 ld_stream_f2(const float *p) {
    float2 v;
    // L1-allocating on purpose: the rows are only 8-byte aligned, so a quad's 32 bytes usually straddle two 32-byte
    // sectors and the neighbouring k-step wants the other half — it must come from L1, not from L2 again
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 32 keep bits starting at flat bit position pos (any alignment); words past the end read as 0
__device__ __forceinline__ uint32_t bit_window(const uint32_t *__restrict__ bits, int64_t words, int64_t pos) {
    const int64_t w = pos >> 5;
    const uint32_t lo = w < words ? __ldg(bits + w) : 0u;
    const uint32_t hi = w + 1 < words ? __ldg(bits + w + 1) : 0u;
    return __funnelshift_r(lo, hi, (uint32_t)(pos & 31));
}

// ------------------------------------------------------------------------------- forward ----
// A warp owns 16 rows: lane (g = lane/4, t = lane%4) loads X[row g][8s+2t, +1] and X[row g+8][...] per
// k-step s.  W sits in shared memory already split and already in B-fragment order, one float4 (big)
// + one float4 (small) per lane per k-step: {b0,b1 of n-tile 0, b0,b1 of n-tile 1}.
constexpr int FW_BATCH = 4;                 // k-steps (8 columns each) per load batch = 32 columns = one keep window
constexpr int FW_PF = 4;                    // L2 prefetch distance in batches (512 B ahead in each of the 16 rows)

struct FwBatch {
    float2 a[FW_BATCH], b[FW_BATCH];        // rows g and g+8
    uint32_t wa, wb;                        // keep windows of the two rows
};

__device__ __forceinline__ void fw_load(FwBatch &q, const float *xa, const float *xb, bool va, bool vb, int col0, int n,
                                        const uint32_t *bits, int64_t words, int64_t pa, int64_t pb) {
#pragma unroll
    for (int j = 0; j < FW_BATCH; j++) {
        const bool in = col0 + 8 * j < n;                       // n is even, so the pair is all in or all out
        q.a[j] = (va && in) ? ld_stream_f2(xa + 8 * j) : make_float2(0.f, 0.f);
        q.b[j] = (vb && in) ? ld_stream_f2(xb + 8 * j) : make_float2(0.f, 0.f);
    }
    if (bits) {
        q.wa = va ? bit_window(bits, words, pa) : 0u;
        q.wb = vb ? bit_window(bits, words, pb) : 0u;
    }
}

__global__ void __launch_bounds__(THREADS, 2) dense_fw16_tc_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                    float *__restrict__ c, int m, int n,
                                                                    const uint32_t *__restrict__ bits, int64_t bit_words,
                                                                    float scale, const float *__restrict__ row_scale, int relu,
                                                                    const Mirror mirror) {
    extern __shared__ float4 sfrag[];       // [KS][32] big, then [KS][32] small
    const int KS = (n + 7) / 8;
    for (int i = threadIdx.x; i < KS * 32; i += THREADS) {
        const int s = i >> 5, l = i & 31, t = l & 3, g = l >> 2;
        const int k0 = 8 * s + 2 * t, k1 = k0 + 1;
        const float v[4] = {k0 < n ? w[k0 * P + g] : 0.f, k1 < n ? w[k1 * P + g] : 0.f,
                            k0 < n ? w[k0 * P + 8 + g] : 0.f, k1 < n ? w[k1 * P + 8 + g] : 0.f};
        uint32_t big[4], small[4];
#pragma unroll
        for (int e = 0; e < 4; e++) split(v[e], big[e], small[e]);
        sfrag[i] = make_float4(__uint_as_float(big[0]), __uint_as_float(big[1]), __uint_as_float(big[2]), __uint_as_float(big[3]));
        sfrag[KS * 32 + i] = make_float4(__uint_as_float(small[0]), __uint_as_float(small[1]), __uint_as_float(small[2]), __uint_as_float(small[3]));
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = lane & 3, g = lane >> 2;
    const int n_tiles = (m + 15) / 16, n_batches = (KS + FW_BATCH - 1) / FW_BATCH;
    for (int tile = blockIdx.x * WARPS + warp; tile < n_tiles; tile += gridDim.x * WARPS) {
        const int ra = tile * 16 + g, rb = ra + 8;
        const bool va = ra < m, vb = rb < m;
        const float *xa = x + (size_t)(va ? ra : 0) * n + 2 * t, *xb = x + (size_t)(vb ? rb : 0) * n + 2 * t;
        const int64_t pa = (int64_t)ra * n, pb = (int64_t)rb * n;
        Acc3 acc0, acc1;
        acc0.zero(); acc1.zero();
        FwBatch cur, nxt;
        fw_load(cur, xa, xb, va, vb, 2 * t, n, bits, bit_words, pa, pb);
#pragma unroll 1
        for (int b = 0; b < n_batches; b++) {
            {   // pull the X lines FW_PF batches ahead (wrapping into this warp's next tile) into L2: the register
                // double buffer then sees L2 latency instead of HBM latency
                int pb_ = b + FW_PF, ptile = tile;
                if (pb_ >= n_batches) { pb_ -= n_batches; ptile += gridDim.x * WARPS; }
                const int pra = ptile * 16 + g, prb = pra + 8, pcol = 32 * pb_ + 2 * t;
                if (pcol < n) {
                    if (pra < m) prefetch_l2(x + (size_t)pra * n + pcol);
                    if (prb < m) prefetch_l2(x + (size_t)prb * n + pcol);
                }
            }
            if (b + 1 < n_batches)                                  // next batch in flight while this one is multiplied
                fw_load(nxt, xa + 32 * (b + 1), xb + 32 * (b + 1), va, vb, 32 * (b + 1) + 2 * t, n, bits, bit_words,
                        pa + 32 * (b + 1), pb + 32 * (b + 1));
            // keep bits of this lane's two columns, k-step j: bits 8j and 8j+1 of the window shifted by 2t
            const uint32_t ka = cur.wa >> (2 * t), kb = cur.wb >> (2 * t);
#pragma unroll
            for (int j = 0; j < FW_BATCH; j++) {
                const int s = b * FW_BATCH + j;
                if (s < KS) {
                    float2 fa = cur.a[j], fb = cur.b[j];
                    if (bits) {                                     // the 1/(1-p) factor is applied once, in the epilogue
                        fa.x = ka & (1u << (8 * j)) ? fa.x : 0.f;
                        fa.y = ka & (2u << (8 * j)) ? fa.y : 0.f;
                        fb.x = kb & (1u << (8 * j)) ? fb.x : 0.f;
                        fb.y = kb & (2u << (8 * j)) ? fb.y : 0.f;
                    }
                    uint32_t ab[4], as[4];
                    split_trunc(fa.x, ab[0], as[0]);      // a0: (row g,   logical k t)   = column 2t
                    split_trunc(fb.x, ab[1], as[1]);      // a1: (row g+8, logical k t)
                    split_trunc(fa.y, ab[2], as[2]);      // a2: (row g,   logical k t+4) = column 2t+1
                    split_trunc(fb.y, ab[3], as[3]);      // a3: (row g+8, logical k t+4)
                    const float4 wb4 = sfrag[s * 32 + lane], ws4 = sfrag[(KS + s) * 32 + lane];
                    mma3(acc0, ab, as, __float_as_uint(wb4.x), __float_as_uint(wb4.y), __float_as_uint(ws4.x), __float_as_uint(ws4.y));
                    mma3(acc1, ab, as, __float_as_uint(wb4.z), __float_as_uint(wb4.w), __float_as_uint(ws4.z), __float_as_uint(ws4.w));
                }
            }
            cur = nxt;
        }
        // D fragment: c0,c1 = (row g, cols 2t,2t+1), c2,c3 = (row g+8, same cols) of each 8-wide n-tile
        float r0[4] = {acc0.total(0), acc0.total(1), acc0.total(2), acc0.total(3)};
        float r1[4] = {acc1.total(0), acc1.total(1), acc1.total(2), acc1.total(3)};
        if (relu) {                                                  // x > 0 ? x : 0 (module.cpp:179-181)
#pragma unroll
            for (int e = 0; e < 4; e++) { r0[e] = r0[e] > 0.f ? r0[e] : 0.f; r1[e] = r1[e] > 0.f ? r1[e] : 0.f; }
        }
        const float post = bits ? scale : 1.f;
        if (va) {
            const float rs = post * (row_scale ? row_scale[ra] : 1.f);
            float *o = c + (size_t)ra * P + 2 * t;
            const float2 lo = make_float2(rs * r0[0], rs * r0[1]), hi = make_float2(rs * r1[0], rs * r1[1]);
            *reinterpret_cast<float2 *>(o) = lo;
            *reinterpret_cast<float2 *>(o + 8) = hi;
            mirror_store(mirror, (size_t)ra * P + 2 * t, lo);
            mirror_store(mirror, (size_t)ra * P + 2 * t + 8, hi);
        }
        if (vb) {
            const float rs = post * (row_scale ? row_scale[rb] : 1.f);
            float *o = c + (size_t)rb * P + 2 * t;
            const float2 lo = make_float2(rs * r0[2], rs * r0[3]), hi = make_float2(rs * r1[2], rs * r1[3]);
            *reinterpret_cast<float2 *>(o) = lo;
            *reinterpret_cast<float2 *>(o + 8) = hi;
            mirror_store(mirror, (size_t)rb * P + 2 * t, lo);
            mirror_store(mirror, (size_t)rb * P + 2 * t + 8, hi);
        }
    }
}

// ------------------------------------------------------------------------------ backward ----
// D^T[16 x n] = G^T[16 x m] * drop(X)[m x n]:  M = 16 hidden units, N = features, K = rows.
// A CTA owns a slab of rows; warp w owns the feature band [BAND*w, BAND*(w+1)) as BW_PAIRS pairs of n-tiles.
// Per k-step (8 rows): A fragment from G (a0 = G[row 2t][g], a1 = G[row 2t][g+8], a2 = G[row 2t+1][g],
// a3 = G[row 2t+1][g+8] — logical k = t / t+4 mapped to the adjacent rows 2t / 2t+1), B fragments from X:
// one 64-bit load X[row][f0 + 2g, +1] feeds n-tile "even" (.x) and n-tile "odd" (.y) of the pair.
constexpr int BW_PAIRS = 5, BAND = 16 * BW_PAIRS;          // 80 features per warp, 640 per CTA
constexpr int BW_PF = 6;                                   // L2 prefetch distance in k-steps (8 rows each)

struct BwStep {
    float2 x0[BW_PAIRS], x1[BW_PAIRS];     // rows k0+2t and k0+2t+1, this lane's column pair of each n-tile pair
    uint32_t w0[3], w1[3];                 // 96 keep bits from the band start of each row
    float a[4];                            // G[r0][g], G[r0][g+8], G[r1][g], G[r1][g+8]
};

__device__ __forceinline__ void bw_load(BwStep &q, const float *__restrict__ x, const float *__restrict__ gmat, int n, int k0, int r_hi,
                                        int f_band, int t, int g, const uint32_t *__restrict__ bits, int64_t words) {
    const int r0 = k0 + 2 * t, r1 = r0 + 1;
    const bool v0 = r0 < r_hi, v1 = r1 < r_hi;
    const float *p0 = x + (size_t)(v0 ? r0 : 0) * n + f_band + 2 * g, *p1 = x + (size_t)(v1 ? r1 : 0) * n + f_band + 2 * g;
#pragma unroll
    for (int p = 0; p < BW_PAIRS; p++) {
        const bool in = f_band + 16 * p + 2 * g < n;
        q.x0[p] = (v0 && in) ? ld_stream_f2(p0 + 16 * p) : make_float2(0.f, 0.f);
        q.x1[p] = (v1 && in) ? ld_stream_f2(p1 + 16 * p) : make_float2(0.f, 0.f);
    }
    if (bits) {
#pragma unroll
        for (int w = 0; w < 3; w++) {
            q.w0[w] = v0 ? bit_window(bits, words, (int64_t)r0 * n + f_band + 32 * w) : 0u;
            q.w1[w] = v1 ? bit_window(bits, words, (int64_t)r1 * n + f_band + 32 * w) : 0u;
        }
    }
    const float *g0 = gmat + (size_t)(v0 ? r0 : 0) * P, *g1 = gmat + (size_t)(v1 ? r1 : 0) * P;
    q.a[0] = v0 ? __ldg(g0 + g) : 0.f;
    q.a[1] = v0 ? __ldg(g0 + g + 8) : 0.f;
    q.a[2] = v1 ? __ldg(g1 + g) : 0.f;
    q.a[3] = v1 ? __ldg(g1 + g + 8) : 0.f;
}

template <bool PF, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) dense_bw16_tc_kernel(const float *__restrict__ x, const float *__restrict__ gmat,
                                                                    float *__restrict__ partials, int m, int n, int rows_per_cta,
                                                                    const uint32_t *__restrict__ bits, int64_t bit_words, float scale) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = lane & 3, g = lane >> 2;
    const int r_lo = blockIdx.x * rows_per_cta, r_hi = min(m, r_lo + rows_per_cta);
    const int f_band = BAND * warp + blockIdx.y * (BAND * WARPS);
    // one accumulator per n-tile: the 10 tiles of the band are 10 independent MMA chains
    float acc[BW_PAIRS][2][4];
#pragma unroll
    for (int p = 0; p < BW_PAIRS; p++)
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[p][h][e] = 0.f;

    if (f_band < n && r_lo < r_hi) {
        BwStep cur, nxt;
        if (PF) bw_load(cur, x, gmat, n, r_lo, r_hi, f_band, t, g, bits, bit_words);
#pragma unroll 1
        for (int k0 = r_lo; k0 < r_hi; k0 += 8) {
            if (!PF) bw_load(cur, x, gmat, n, k0, r_hi, f_band, t, g, bits, bit_words);
            {   // L2 prefetch of this warp's band BW_PF k-steps ahead (the band is 320 B per row: lanes g cover it in 64-B steps)
                const int pr = k0 + 8 * BW_PF + 2 * t + (g >> 2);
                const int pf = f_band + 32 * (g & 3) * 1;
                if (pr < r_hi && pf < n && (g & 3) * 32 < BAND) prefetch_l2(x + (size_t)pr * n + pf);
            }
            if (PF && k0 + 8 < r_hi) bw_load(nxt, x, gmat, n, k0 + 8, r_hi, f_band, t, g, bits, bit_words);   // next 8 rows in flight
            uint32_t ab[4], as[4];
#pragma unroll
            for (int e = 0; e < 4; e++) split_trunc(cur.a[e], ab[e], as[e]);
#pragma unroll
            for (int p = 0; p < BW_PAIRS; p++) {
                float2 f0 = cur.x0[p], f1 = cur.x1[p];
                if (bits) {                                          // the 1/(1-p) factor is applied to the partial sums
                    // bit 16p + 2g of the 96-bit window; 2g is even, so both bits sit in the same word
                    const int q = (16 * p) >> 5, s = (16 * p) & 31;
                    const uint32_t k0w = cur.w0[q] >> (2 * g), k1w = cur.w1[q] >> (2 * g);
                    // 16p mod 32 is 0 or 16 and 2g <= 14: the pair never straddles a word
                    f0.x = k0w & (1u << s) ? f0.x : 0.f;
                    f0.y = k0w & (2u << s) ? f0.y : 0.f;
                    f1.x = k1w & (1u << s) ? f1.x : 0.f;
                    f1.y = k1w & (2u << s) ? f1.y : 0.f;
                }
                uint32_t bb[4], bs[4];
                split_trunc(f0.x, bb[0], bs[0]);      // even n-tile: b0 (logical k t   = row 2t)
                split_trunc(f1.x, bb[1], bs[1]);      //              b1 (logical k t+4 = row 2t+1)
                split_trunc(f0.y, bb[2], bs[2]);      // odd n-tile
                split_trunc(f1.y, bb[3], bs[3]);
                mma(acc[p][0], as[0], as[1], as[2], as[3], bb[0], bb[1]);
                mma(acc[p][1], as[0], as[1], as[2], as[3], bb[2], bb[3]);
                mma(acc[p][0], ab[0], ab[1], ab[2], ab[3], bs[0], bs[1]);
                mma(acc[p][1], ab[0], ab[1], ab[2], ab[3], bs[2], bs[3]);
                mma(acc[p][0], ab[0], ab[1], ab[2], ab[3], bb[0], bb[1]);
                mma(acc[p][1], ab[0], ab[1], ab[2], ab[3], bb[2], bb[3]);
            }
            if (PF) cur = nxt;
        }
    }
    // D fragment of an n-tile: c0 = (h g, n 2t), c1 = (h g, n 2t+1), c2 = (h g+8, n 2t), c3 = (h g+8, n 2t+1);
    // logical column n of the even tile is feature f0 + 2n, of the odd tile f0 + 2n + 1.
    const float post = bits ? scale : 1.f;
    float *out = partials + (size_t)blockIdx.x * n * P;
#pragma unroll
    for (int p = 0; p < BW_PAIRS; p++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int fa = f_band + 16 * p + 2 * (2 * t) + h, fb = f_band + 16 * p + 2 * (2 * t + 1) + h;
            if (fa < n) { out[(size_t)fa * P + g] = post * acc[p][h][0]; out[(size_t)fa * P + g + 8] = post * acc[p][h][2]; }
            if (fb < n) { out[(size_t)fb * P + g] = post * acc[p][h][1]; out[(size_t)fb * P + g + 8] = post * acc[p][h][3]; }
        }
}

// sums the per-CTA partials in a fixed order (deterministic)
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float *__restrict__ partials, float *__restrict__ out, int elems, int parts) {
    reduce_parts_block(partials, out, elems, parts);
}

}  // namespace gcnk_tc

// -------------------------------------------------------------------------------------------------
// host-side launchers, called from feature.cu
namespace gcnk {

bool dense_tc_supported(int n, int p) { return p == gcnk_tc::P && n % 2 == 0 && n >= 8; }

int dense_fw16_tc(const float *x, const float *w, float *c, int m, int n, const uint32_t *bits, int64_t nnz, float scale,
                  const float *row_scale, int relu, cudaStream_t st) {
    using namespace gcnk_tc;
    const int KS = (n + 7) / 8;
    const size_t smem = sizeof(float4) * 2 * (size_t)KS * 32;
    if (smem > 100 * 1024) return GCNK_EUNSUPPORTED;                   // two CTAs per SM must fit
    static bool attr_set[64] = {false};
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev]) {
        GCNK_CUDA(cudaFuncSetAttribute(dense_fw16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set[dev] = true;
    }
    const int n_tiles = (m + 15) / 16;
    const Mirror mir = take_mirror(c);
    // (A cp.async variant — 8-byte copies into a 4-stage shared-memory ring per warp, 16 warps/SM — was measured at
    // 293 us per pass against 176 us for the register double buffer below: with rows that are only 8-byte aligned the
    // copies cannot be wider than 8 bytes, and LDGSTS.64 costs more issue slots than it frees.  A 16-byte-aligned
    // re-packed copy of X (row pitch 608 floats) would allow 16-byte cp.async / TMA tiles; not done yet.)
    const int grid = std::max(1, std::min(sm_count() * 2, (n_tiles + WARPS - 1) / WARPS));
    dense_fw16_tc_kernel<<<grid, THREADS, smem, st>>>(x, w, c, m, n, bits, (nnz + 31) / 32, scale, row_scale, relu, mir);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

size_t dense_bw16_tc_parts(int m) { return (size_t)std::max(1, std::min(sm_count() * 2, (m + 63) / 64)); }

int dense_bw16_tc(const float *x, const float *g, float *b_grad, float *partials, int m, int n, const uint32_t *bits, int64_t nnz,
                  float scale, cudaStream_t st) {
    using namespace gcnk_tc;
    const int f_groups = (n + BAND * WARPS - 1) / (BAND * WARPS);
    int ctas = (int)dense_bw16_tc_parts(m);
    if (f_groups > 1) ctas = std::max(1, ctas / f_groups);
    int rows_per_cta = (m + ctas - 1) / ctas;
    rows_per_cta = (rows_per_cta + 7) / 8 * 8;
    const int parts = (m + rows_per_cta - 1) / rows_per_cta;
    dim3 grid(parts, f_groups, 1);
    // (variants without the register prefetch at 3 and 4 CTAs/SM — 80 / 64 registers, some spills — measured the same
    // 0.28-0.31 ms per pass as this one: occupancy is not what limits this kernel)
    dense_bw16_tc_kernel<true, 2><<<grid, THREADS, 0, st>>>(x, g, partials, m, n, rows_per_cta, bits, (nnz + 31) / 32, scale);
    GCNK_LAUNCHED();
    const int elems = n * P;
    reduce_parts_kernel<<<(elems + 31) / 32, 256, 0, st>>>(partials, b_grad, elems, parts);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // namespace gcnk
