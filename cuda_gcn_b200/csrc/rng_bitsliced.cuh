// rng_bitsliced.cuh — the per-thread part of the bit-sliced xorshift128+ keep-bit generator (rng.cu), written so that the
// same code compiles for the device and for the host (tests/test_host_cpu.py runs it under g++ against the scalar stream).
//
// The scalar generator (rand.cpp:17-28 of the reference) spends ~24 instructions per draw on 64-bit shifts and XORs.
// Bit-sliced, one thread runs 32 independent streams at once: register i holds bit i of the state of all 32 streams, so
// a shift is a renaming of registers (free) and every XOR works on 32 streams.  Per step of all 32 streams:
//     152 LOP3 for the state map, 31 x 3 for the 31-bit sum s0 + s1 and the comparison with the threshold (both
//     LSB-first, so they fuse into one pass over the bits),
// i.e. ~7.7 instructions per draw.  The 32 streams of a thread are consecutive runs of 2^LS draws; their start states come
// from one chain of 31 jumps by 2^LS draws (128 x 128 bit matrix applied through 32 nibble tables of 16 entries); the keep
// words come out transposed (bit k of word j = stream k, step j) and are turned into draw order by a 32 x 32 bit transpose
// every 32 steps.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define GCNK_BS_HD __host__ __device__ __forceinline__
#else
#define GCNK_BS_HD inline
#endif

namespace gcnk_bs {

struct State128 { uint64_t lo, hi; };                       // lo = s0, hi = s1 (as U128 in rng.cu)
struct alignas(16) Entry { uint32_t w[4]; };                // one 128-bit table entry: lo low/high word, hi low/high word

constexpr int NIB_ENTRIES = 32 * 16;                        // nibble tables of one jump matrix: [nibble 0..31][value 0..15]

// Host side: the nibble tables of the matrix whose 128 columns are cols[j] = M * e_j.
inline void build_nibble_tables(const State128 *cols, Entry *tab) {
    for (int n = 0; n < 32; n++)
        for (int v = 0; v < 16; v++) {
            uint64_t lo = 0, hi = 0;
            for (int i = 0; i < 4; i++)
                if (v >> i & 1) { lo ^= cols[4 * n + i].lo; hi ^= cols[4 * n + i].hi; }
            tab[n * 16 + v] = Entry{{(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)}};
        }
}

GCNK_BS_HD Entry load_entry(const Entry *p) {
#ifdef __CUDA_ARCH__
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));   // 8 KB per matrix: stays in L1
    return Entry{{v.x, v.y, v.z, v.w}};
#else
    return *p;
#endif
}

// s -> M * s through the nibble tables of M
GCNK_BS_HD void apply_nibbles(const Entry *tab, uint32_t (&s)[4]) {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const Entry e = load_entry(tab + (w * 8 + q) * 16 + ((s[w] >> (4 * q)) & 15u));
            r0 ^= e.w[0]; r1 ^= e.w[1]; r2 ^= e.w[2]; r3 ^= e.w[3];
        }
    }
    s[0] = r0; s[1] = r1; s[2] = r2; s[3] = r3;
}

// 32 x 32 bit transpose in registers: afterwards bit k of A[i] is what bit i of A[k] was.
GCNK_BS_HD void transpose32(uint32_t (&A)[32]) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if (k & j) continue;
            const uint32_t t = ((A[k] >> j) ^ A[k + j]) & m;     // bits j.. of row k  <->  bits 0.. of row k + j
            A[k] ^= t << j;
            A[k + j] ^= t;
        }
    }
}

// One step of all 32 streams.  a = bits of s0, b = bits of s1; afterwards a holds the NEW s1 and b the new s0 (the old
// s1), so the caller alternates step(a, b), step(b, a).  Returns the keep word: bit k = (draw of stream k >= threshold),
// draw = (s0 + s1) & 0x7fffffff of the new state (rand.cpp:26).
// HALF: threshold == 2^30 (dropout 0.5, the reference's default): the draw is kept iff bit 30 of the sum is set, so the
// lower 30 bits only have to deliver their carry (1 instruction per bit instead of 3).
template <bool HALF>
GCNK_BS_HD uint32_t step32(uint32_t (&a)[64], const uint32_t (&b)[64], uint32_t threshold) {
#pragma unroll
    for (int i = 63; i >= 23; i--) a[i] ^= a[i - 23];             // t ^= t << 23
#pragma unroll
    for (int i = 0; i < 47; i++) a[i] ^= a[i + 17];               // t ^= t >> 17
#pragma unroll
    for (int i = 0; i < 64; i++) a[i] ^= i + 26 < 64 ? b[i] ^ b[i + 26] : b[i];   // t ^= s ^ (s >> 26)
    uint32_t carry = 0, ge = 0xffffffffu;                          // equal so far counts as >=
    if (HALF) {
#pragma unroll
        for (int i = 0; i < 30; i++) carry = (a[i] & b[i]) | (carry & (a[i] ^ b[i]));
        return a[30] ^ b[30] ^ carry;
    }
#pragma unroll
    for (int i = 0; i < 31; i++) {
        const uint32_t d = a[i] ^ b[i] ^ carry;                    // bit i of the sum
        if (i < 30) carry = (a[i] & b[i]) | (carry & (a[i] ^ b[i]));
        const uint32_t t = 0u - ((threshold >> i) & 1u);           // all ones where the threshold has bit i set (uniform, loop-invariant)
        ge = (t & (d & ge)) | (~t & (d | ge));
    }
    return ge;
}

// All draws of one thread: 32 streams of 2^LS draws starting at `base`, keep bits in draw order at out[0 ..], only the first
// `left` draws exist (bits beyond stay 0 in a partial word, words beyond are not written).
// jump_tab = nibble tables of M^(2^LS).
template <int LS, bool HALF>
GCNK_BS_HD void generate(State128 base, const Entry *jump_tab, int threshold, uint32_t *out, int64_t left) {
    constexpr int L = 1 << LS, WORDS = L / 32;
    static_assert(LS >= 5, "a stream is at least one keep word");
    const uint32_t thr = threshold > 0 ? (uint32_t)threshold : 0u; // draws are non-negative
    uint32_t a[64], b[64];
    {
        // start states of the 32 streams: a chain of jumps by 2^LS draws, eight per trip (the shifting keeps the
        // indices static without unrolling all 31 table walks)
        uint32_t w0[32], w1[32], w2[32], w3[32];
#pragma unroll
        for (int i = 0; i < 32; i++) w0[i] = w1[i] = w2[i] = w3[i] = 0;
        uint32_t cur[4] = {(uint32_t)base.lo, (uint32_t)(base.lo >> 32), (uint32_t)base.hi, (uint32_t)(base.hi >> 32)};
#pragma unroll 1
        for (int grp = 0; grp < 4; grp++) {
#pragma unroll
            for (int i = 0; i < 24; i++) { w0[i] = w0[i + 8]; w1[i] = w1[i + 8]; w2[i] = w2[i + 8]; w3[i] = w3[i + 8]; }
#pragma unroll
            for (int kk = 0; kk < 8; kk++) {
                w0[24 + kk] = cur[0]; w1[24 + kk] = cur[1]; w2[24 + kk] = cur[2]; w3[24 + kk] = cur[3];
                if ((int64_t)(grp * 8 + kk + 1) * L < left) apply_nibbles(jump_tab, cur);   // streams that do not exist keep a stale state
            }
        }
        transpose32(w0); transpose32(w1); transpose32(w2); transpose32(w3);
#pragma unroll
        for (int i = 0; i < 32; i++) { a[i] = w0[i]; a[32 + i] = w1[i]; b[i] = w2[i]; b[32 + i] = w3[i]; }
    }
    const int64_t max_steps = left < L ? left : L;                 // fewer than L only when just stream 0 exists
    const bool whole = left >= (int64_t)32 * L;                    // every word of every stream exists: plain stores
    for (int m = 0; m < WORDS && (int64_t)m * 32 < max_steps; m++) {
        uint32_t kw[32];
#pragma unroll
        for (int i = 0; i < 32; i++) kw[i] = 0;
#pragma unroll 1
        for (int grp = 0; grp < 4; grp++) {
#pragma unroll
            for (int i = 0; i < 24; i++) kw[i] = kw[i + 8];        // oldest step at the lowest index
#pragma unroll
            for (int jj = 0; jj < 8; jj += 2) {
                kw[24 + jj] = step32<HALF>(a, b, thr);
                kw[25 + jj] = step32<HALF>(b, a, thr);
            }
        }
        transpose32(kw);                                           // kw[k] = 32 consecutive keep bits of stream k
        if (whole) {
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * WORDS + m] = kw[k];
        } else {
#pragma unroll 1
            for (int k = 0; k < 32; k++) {
                uint32_t word = 0;
#pragma unroll
                for (int i = 0; i < 32; i++) word = i == k ? kw[i] : word;   // (static indices: kw stays in registers)
                const int64_t pos = (int64_t)k * L + 32 * m;      // first draw of this word, relative to the thread's first
                if (pos < left) out[(size_t)k * WORDS + m] = left - pos >= 32 ? word : (word & ((1u << (int)(left - pos)) - 1u));
            }
        }
    }
}

}  // namespace gcnk_bs
