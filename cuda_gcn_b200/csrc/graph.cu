// graph.cu — GraphSum, the hot path.  Replaces cuda_GraphSum_forward/backward_kernel
// (reference src/cuda/cuda_kernel.cu:126-162; CPU semantics src/seq/module.cpp:83-119).
//
// Design (B200-first, not a translation):
//   * gcnk_graph_create() does once per graph what the reference redoes per edge per launch: the
//     degree vector d^-1/2, and a STATIC degree-balanced schedule (longest-processing-time bins of
//     rows per warp, rows above a degree threshold get a whole CTA).  Static => every sum has a fixed
//     order => results are reproducible run to run and identical for 1 and N ranks.
//   * the gather source is pre-scaled by d^-1/2 of its own row (by the producing kernel's epilogue),
//     so an edge costs one index (streamed, coalesced, L1-bypassing) and one vectorised row gather;
//     no sqrtf, no division, no second random indptr read per edge.
//   * at width 16 (and 12) the lanes that share a row fetch their indices as one aligned int4 and issue the
//     chunk's four row reads back to back: the kernel is bound by the L1TEX LSU data pipe (one wavefront per
//     gathered 64-byte row), and warp shuffles run on that pipe too.
//   * each warp keeps the row sum in registers and WRITES the row once (no global +=, no memset).
//   * the epilogue fuses the degree normalisation with ReLU + Dropout (+ mask) forward, or with the
//     Dropout/ReLU backward mask, and with the pre-scale for the next gather.
//
// Roofline: HBM traffic is 4*nnz (indices) + 8*n*dim; the edge gathers (dim*4 bytes each) are served
// by L2, where the [n x dim] source is resident (14.9 MB at Reddit shape, dim 16); on chip, one L1TEX
// wavefront per edge per SM clock is the floor (395 us at Reddit shape; measured 460 us).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <queue>
#include <vector>

#include "common.cuh"

using namespace gcnk;

struct gcnk_graph {
    const int *indptr = nullptr, *indices = nullptr;
    int n = 0, n_cols = 0;
    int64_t nnz = 0;
    float *dinv = nullptr;            // [n] d^-1/2 of the local rows
    const float *dinv_cols = nullptr; // [n_cols] d^-1/2 of the gather-source rows (== dinv when n_cols == n)
    int *heavy_rows = nullptr; int n_heavy = 0;
    int *bin_ptr = nullptr, *bin_rows = nullptr; int n_bins = 0;   // n_bins is a multiple of WARPS
    int max_degree = 0, symmetric = 0;
    int idx4_ok = 0;                  // `indices` is 16-byte aligned and readable up to its length rounded up to 4 entries
    float *scratch = nullptr; size_t scratch_elems = 0;            // pre-scaled copy for gcnk_graphsum
    // views (gcnk_graph_create_view): dinv/dinv_cols are borrowed from `base`; a column-filtered view owns its CSR
    const gcnk_graph *base = nullptr;
    int *own_indptr = nullptr, *own_indices = nullptr;
    int n_rows_scheduled = 0;
    // rotated order (gcnk_graph_rotate, row-partitioned runs): every row lists the columns this rank owns first, then the
    // columns of higher ranks, then those of lower ranks; seg1[s] / seg2[s] are the two boundaries (absolute offsets)
    int rot_lo = -1, rot_hi = -1;
    int *seg1 = nullptr, *seg2 = nullptr;
    bool seg_owned = false;
};

namespace {

constexpr int THREADS = 256, WARPS = THREADS / 32;
constexpr int HEAVY_DEGREE = 2048;     // rows above this get a whole CTA (less for small graphs, see build_schedule)
constexpr int ROW_OVERHEAD = 24;       // per-row cost in edge-equivalents for the bin balance

enum Mode { MODE_PLAIN = 0, MODE_RELU_DROP = 1, MODE_MASK = 2, MODE_RAW = 3 };   // RAW: the unscaled row sum (a partial result)

struct GatherArgs {
    const int *indptr, *indices;
    const float *in;
    float *out;
    const float *dinv;
    const int *heavy_rows, *bin_ptr, *bin_rows;
    int n_heavy, dim, mode, mask_stride;   // mask_stride: bits per row of the written/read mask
    const uint32_t *drop_bits;             // flat: bit s*dim+j (may be NULL)
    uint32_t *mask_out;                    // MODE_RELU_DROP (may be NULL)
    const uint32_t *mask_in;               // MODE_MASK
    float scale;
    // row-partitioned runs: flags[r] >= wait_value once rank r's rows of the gather source have landed in this GPU's
    // copy (gcnk_peer_push_signal); every CTA checks them before its first row read.  NULL: no wait.
    const int *wait_flags;
    int wait_n, wait_skip, wait_value;
    int *wait_err;
    long long wait_limit;
    // accumulate: the row sum starts from init[s, :] (the raw partial sum an earlier launch over other columns left there)
    const float *init;
    // fused exchange (xgather_kernel): the first x_ctas CTAs copy this rank's rows of the source to the peers and publish
    // x_value in their flag slots; the others aggregate in rotated order, waiting for a peer's rows only when they get there
    const float4 *x_src;
    float4 *x_dst[7];
    int *x_flag[7];
    size_t x_vec;
    unsigned *x_counter;
    int x_peers, x_ctas, x_value, x_rank, x_world;
    const int *seg1, *seg2;
};

// Registered by gcnk_gather_wait_next for the next gather launched by this thread.
struct GatherWait { const int *flags; int n, skip, value; int *err; };
thread_local GatherWait t_wait = {nullptr, 0, -1, 0, nullptr};
thread_local const float *t_init = nullptr;     // gcnk_gather_init_next
struct GatherXchg { const float *src; float *dst[7]; int *flag[7]; size_t n_floats; unsigned *counter; int peers, value, rank, world; const int *wait_flags; int *err; bool armed; };
thread_local GatherXchg t_xchg = {};            // gcnk_gather_exchange_next

// All rows of the source that other ranks produce must be in place before any of them is read: thread r of every CTA
// polls rank r's flag (an acquire load at system scope: the peer wrote the rows, fenced, then the flag), the CTA
// barrier orders everybody else's reads after it.  Nothing of the source has been touched by this kernel before, so
// no stale line can sit in L1.  A peer that never arrives raises *err after wait_limit clocks instead of hanging.
__device__ __forceinline__ void wait_for_peers(const GatherArgs &a) {
    if (!a.wait_flags) return;
    const int r = threadIdx.x;
    if (r < a.wait_n && r != a.wait_skip) {
        const int *f = a.wait_flags + r;
        const long long t0 = clock64();
        for (;;) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v >= a.wait_value) break;
            if (clock64() - t0 > a.wait_limit) { *a.wait_err = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

__host__ __device__ inline int mask_stride_bits(int dim) { return dim <= 8 ? 8 : dim <= 16 ? 16 : (dim + 31) / 32 * 32; }

// row gathers go through L1 (LDG.CONSTANT): bypassing it (L1::no_allocate) was measured 16 % slower even at a 1 % hit rate
#ifndef GCNK_GATHER_LD4
#define GCNK_GATHER_LD4(p) __ldg(p)
#endif

template <int VEC> struct Acc;
template <> struct Acc<4> {
    float4 v;
    __device__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ void load_add(const float *p) {
        const float4 x = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(p));
        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
    }
    __device__ void add(const Acc &o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
    __device__ void load(const float *p) { v = *reinterpret_cast<const float4 *>(p); }
    __device__ void shfl_xor_add(int off) {
        v.x += __shfl_xor_sync(FULL, v.x, off); v.y += __shfl_xor_sync(FULL, v.y, off);
        v.z += __shfl_xor_sync(FULL, v.z, off); v.w += __shfl_xor_sync(FULL, v.w, off);
    }
    __device__ float get(int i) const { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
    __device__ void set(int i, float f) { if (i == 0) v.x = f; else if (i == 1) v.y = f; else if (i == 2) v.z = f; else v.w = f; }
};
template <> struct Acc<1> {
    float v;
    __device__ void zero() { v = 0.f; }
    __device__ void load_add(const float *p) { v += __ldg(p); }
    __device__ void add(const Acc &o) { v += o.v; }
    __device__ void load(const float *p) { v = *p; }
    __device__ void shfl_xor_add(int off) { v += __shfl_xor_sync(FULL, v, off); }
    __device__ float get(int) const { return v; }
    __device__ void set(int, float f) { v = f; }
};

// One warp sums the gather-source rows of edges [beg,end), taking 32-edge chunks beg+32*(first+k*step).
// Lane layout: LPR lanes per source row (q = lane % LPR covers elements (q + LPR*a)*VEC..+VEC),
// G = 32/LPR edges in flight per gather instruction (g = lane / LPR).
template <int VEC, int LPR, int NACC, bool EXACT>
__device__ __forceinline__ void accumulate(Acc<VEC> (&acc)[NACC], const GatherArgs &a, int beg, int end, int first,
                                           int step, int lane) {
    constexpr int G = 32 / LPR;
    const int g = lane / LPR, q = lane % LPR;
    const int dim = a.dim;
    int e = beg + 32 * first;
    int idx = (e + lane < end) ? ld_stream_i32(a.indices + e + lane) : -1;
    while (e < end) {
        const int e_next = e + 32 * step;
        // software pipeline: the next chunk's indices are in flight while this chunk's rows are gathered
        const int idx_next = (e_next + lane < end) ? ld_stream_i32(a.indices + e_next + lane) : -1;
#pragma unroll
        for (int j = 0; j < LPR; j++) {
            const int d = __shfl_sync(FULL, idx, j * G + g);
            if (d >= 0) {
                const float *row = a.in + (size_t)d * dim;
#pragma unroll
                for (int t = 0; t < NACC; t++) {
                    const int u = (q + LPR * t) * VEC;
                    if (EXACT || u < dim) acc[t].load_add(row + u);   // EXACT: dim == VEC*LPR*NACC
                }
            }
        }
        idx = idx_next;
        e = e_next;
    }
}

// Index fetch for the LPR == 4 layout (dim 12 and 16 — 16 is the hidden width of the benchmark): the four lanes of group g read
// the SAME aligned int4 = the group's four edges of this 32-edge chunk (one 128-byte wavefront per warp, the hardware
// broadcasts inside the group).  The coalesced-load + 4 x SHFL version above spends 5 wavefronts of the L1TEX data pipe
// per chunk on indices — SHFL runs on that pipe too, and the pipe is what bounds this kernel (86 % busy, 14 % of it
// shuffles: profiles/r01h).  Chunks start at the row's begin rounded down to 4 entries; entries of the first and last
// chunk outside [beg,end) are other rows' entries (or the array's padding, see idx4_ok) and are skipped.
__device__ __forceinline__ int4 load_idx4(const int *p) {
    int4 r;
    asm("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// A chunk wholly inside the row: four unconditional row reads per group.  BATCH issues them back to back (16 registers
// of rows in flight: the 40-register build); otherwise the compiler staggers reads and adds to stay within 32.
template <bool BATCH>
__device__ __forceinline__ void gather4_full(Acc<4> &acc, const float *in_q, int dim, const int4 &idx) {
    if constexpr (BATCH) {
        const float4 x0 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)idx.x * dim));
        const float4 x1 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)idx.y * dim));
        const float4 x2 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)idx.z * dim));
        const float4 x3 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)idx.w * dim));
        acc.v.x += x0.x; acc.v.y += x0.y; acc.v.z += x0.z; acc.v.w += x0.w;
        acc.v.x += x1.x; acc.v.y += x1.y; acc.v.z += x1.z; acc.v.w += x1.w;
        acc.v.x += x2.x; acc.v.y += x2.y; acc.v.z += x2.z; acc.v.w += x2.w;
        acc.v.x += x3.x; acc.v.y += x3.y; acc.v.z += x3.z; acc.v.w += x3.w;
    } else {
        acc.load_add(in_q + (size_t)(unsigned)idx.x * dim);
        acc.load_add(in_q + (size_t)(unsigned)idx.y * dim);
        acc.load_add(in_q + (size_t)(unsigned)idx.z * dim);
        acc.load_add(in_q + (size_t)(unsigned)idx.w * dim);
    }
}
// The first / last chunk of a row: the group's entries j with lo <= j < rem (same order of additions as above).
__device__ __forceinline__ void gather4_edge(Acc<4> &acc, const float *in_q, int dim, const int4 &idx, int lo, int rem) {
    if (lo <= 0 && rem > 0) acc.load_add(in_q + (size_t)(unsigned)idx.x * dim);
    if (lo <= 1 && rem > 1) acc.load_add(in_q + (size_t)(unsigned)idx.y * dim);
    if (lo <= 2 && rem > 2) acc.load_add(in_q + (size_t)(unsigned)idx.z * dim);
    if (lo <= 3 && rem > 3) acc.load_add(in_q + (size_t)(unsigned)idx.w * dim);
}

// Two chunks wholly inside the row: eight row reads in flight (the 64-register build), added in the same order.
__device__ __forceinline__ void gather8_full(Acc<4> &acc, const float *in_q, int dim, const int4 &ia, const int4 &ib) {
    const float4 x0 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ia.x * dim));
    const float4 x1 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ia.y * dim));
    const float4 x2 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ia.z * dim));
    const float4 x3 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ia.w * dim));
    const float4 x4 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ib.x * dim));
    const float4 x5 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ib.y * dim));
    const float4 x6 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ib.z * dim));
    const float4 x7 = GCNK_GATHER_LD4(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)ib.w * dim));
    acc.v.x += x0.x; acc.v.y += x0.y; acc.v.z += x0.z; acc.v.w += x0.w;
    acc.v.x += x1.x; acc.v.y += x1.y; acc.v.z += x1.z; acc.v.w += x1.w;
    acc.v.x += x2.x; acc.v.y += x2.y; acc.v.z += x2.z; acc.v.w += x2.w;
    acc.v.x += x3.x; acc.v.y += x3.y; acc.v.z += x3.z; acc.v.w += x3.w;
    acc.v.x += x4.x; acc.v.y += x4.y; acc.v.z += x4.z; acc.v.w += x4.w;
    acc.v.x += x5.x; acc.v.y += x5.y; acc.v.z += x5.z; acc.v.w += x5.w;
    acc.v.x += x6.x; acc.v.y += x6.y; acc.v.z += x6.z; acc.v.w += x6.w;
    acc.v.x += x7.x; acc.v.y += x7.y; acc.v.z += x7.z; acc.v.w += x7.w;
}

// STEP: chunks between two of this warp's chunks (1: the warp owns the row, WARPS: a CTA shares it).
template <bool EXACT, int STEP, int BATCH>   // BATCH: row reads in flight together: 0 = compiler's choice, 4, 8
__device__ __forceinline__ void accumulate_idx4(Acc<4> &acc, const GatherArgs &a, int beg, int end, int first, int lane) {
    constexpr int STRIDE = 32 * STEP;
    const int g4 = (lane >> 2) * 4, q = lane & 3;
    const int dim = EXACT ? 16 : a.dim;                    // a compile-time row pitch when the row is exactly 16 wide
    const int base = (beg & ~3) + 32 * first;              // warp-uniform start of this warp's first chunk
    int left = end - base;                                 // warp-uniform: entries from the chunk start to the row's end
    if (left <= 0) return;
    const int *ip = a.indices + base + g4;                 // this group's four entries of the current chunk
    const float *in_q = a.in + q * 4;
    // a group whose four entries start at or after the row's end does not read (this keeps every read inside the array
    // length rounded up to 4 entries, which is what idx4_ok has verified)
    int4 idx = make_int4(0, 0, 0, 0);
    if (left > g4) idx = load_idx4(ip);
    if (!EXACT && q * 4 >= dim) return;                    // lanes beyond the row width
    if constexpr (BATCH == 8) {
        int4 idx_b = idx;
        if (left - STRIDE > g4) idx_b = load_idx4(ip + STRIDE);
        bool head = base < beg;                            // only this warp's first chunk can start before the row
        for (;;) {
            int4 next_a = idx, next_b = idx_b;
            if (left - 2 * STRIDE > g4) next_a = load_idx4(ip + 2 * STRIDE);   // in flight during the row gathers
            if (left - 3 * STRIDE > g4) next_b = load_idx4(ip + 3 * STRIDE);
            if (!head && left >= STRIDE + 32) {
                gather8_full(acc, in_q, dim, idx, idx_b);
            } else {
                if (!head && left >= 32) gather4_full<true>(acc, in_q, dim, idx);
                else gather4_edge(acc, in_q, dim, idx, head ? beg - base - g4 : 0, left - g4);
                if (left > STRIDE) {
                    if (left - STRIDE >= 32) gather4_full<true>(acc, in_q, dim, idx_b);
                    else gather4_edge(acc, in_q, dim, idx_b, 0, left - STRIDE - g4);
                }
            }
            head = false;
            if (left <= 2 * STRIDE) break;
            ip += 2 * STRIDE;
            left -= 2 * STRIDE;
            idx = next_a;
            idx_b = next_b;
        }
    } else if constexpr (BATCH == 4) {
        int4 idx_next = idx;
        if (left - STRIDE > g4) idx_next = load_idx4(ip + STRIDE);   // in flight during this chunk's row gathers
        if (base >= beg && left >= 32) gather4_full<true>(acc, in_q, dim, idx);
        else gather4_edge(acc, in_q, dim, idx, beg - base - g4, left - g4);
#pragma unroll 1
        while (left > STRIDE) {
            ip += STRIDE;
            left -= STRIDE;
            idx = idx_next;
            if (left - STRIDE > g4) idx_next = load_idx4(ip + STRIDE);
            if (left >= 32) gather4_full<true>(acc, in_q, dim, idx);
            else gather4_edge(acc, in_q, dim, idx, 0, left - g4);
        }
    } else {
        // 32 registers leave no room for a second index vector: the next chunk's indices are requested when this
        // chunk's rows have been added (the other 63 warps of the SM cover the round trip)
        if (base >= beg && left >= 32) gather4_full<false>(acc, in_q, dim, idx);
        else gather4_edge(acc, in_q, dim, idx, beg - base - g4, left - g4);
#pragma unroll 1
        while (left > STRIDE) {
            ip += STRIDE;
            left -= STRIDE;
            if (left > g4) idx = load_idx4(ip);
            if (left >= 32) gather4_full<false>(acc, in_q, dim, idx);
            else gather4_edge(acc, in_q, dim, idx, 0, left - g4);
        }
    }
}

template <int VEC, int LPR, int NACC>
__device__ __forceinline__ void reduce_groups(Acc<VEC> (&acc)[NACC]) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
        for (int t = 0; t < NACC; t++) acc[t].shfl_xor_add(off);
}

// Row epilogue.  Called by the whole warp (uses warp-wide reductions for the mask words); the row sum
// is valid in lanes < LPR.
template <int VEC, int LPR, int NACC>
__device__ __forceinline__ void epilogue(Acc<VEC> (&acc)[NACC], const GatherArgs &a, int s, int lane) {
    const int q = lane % LPR, dim = a.dim;
    const bool owner = lane < LPR;
    const float di = a.dinv[s];
    float *orow = a.out + (size_t)s * dim;
    if (a.mode == MODE_PLAIN || a.mode == MODE_RAW) {
        if (owner) {
            const float di = a.mode == MODE_RAW ? 1.0f : a.dinv[s];
#pragma unroll
            for (int t = 0; t < NACC; t++) {
                const int u = (q + LPR * t) * VEC;
                if (u < dim) {
                    if constexpr (VEC == 4) {
                        const float4 o = make_float4(di * acc[t].v.x, di * acc[t].v.y, di * acc[t].v.z, di * acc[t].v.w);
                        *reinterpret_cast<float4 *>(orow + u) = o;
                    } else {
                        orow[u] = di * acc[t].v;
                    }
                }
            }
        }
        return;
    }
    const int words = (dim + 31) / 32;
    const size_t mbase = (size_t)s * a.mask_stride;   // bit offset of this row in the padded mask
    for (int w = 0; w < words; w++) {
        uint32_t bits = 0;
        // which of this lane's elements fall into mask word w
#pragma unroll
        for (int t = 0; t < NACC; t++) {
            const int u = (q + LPR * t) * VEC;
            if (owner && u < dim && (u >> 5) == w) {
                Acc<VEC> o;
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    const int j = u + v;
                    const float x = di * acc[t].get(v);
                    bool pass;
                    if (a.mode == MODE_RELU_DROP) {
                        bool keep = true;
                        if (a.drop_bits) {
                            const size_t b = (size_t)s * dim + j;
                            keep = (a.drop_bits[b >> 5] >> (b & 31)) & 1u;
                        }
                        pass = (x > 0.f) && keep;
                    } else {
                        const size_t b = mbase + j;
                        pass = (a.mask_in[b >> 5] >> (b & 31)) & 1u;
                    }
                    o.set(v, pass ? di * (x * a.scale) : 0.f);
                    bits |= (uint32_t)pass << (j & 31);
                }
                if constexpr (VEC == 4) *reinterpret_cast<float4 *>(orow + u) = o.v;
                else orow[u] = o.v;
            }
        }
        if (a.mode == MODE_RELU_DROP && a.mask_out) {
            bits = __reduce_or_sync(FULL, bits);
            if (lane == 0) {
                if (a.mask_stride == 8) reinterpret_cast<uint8_t *>(a.mask_out)[s] = (uint8_t)bits;
                else if (a.mask_stride == 16) reinterpret_cast<uint16_t *>(a.mask_out)[s] = (uint16_t)bits;
                else a.mask_out[(mbase >> 5) + w] = bits;
            }
        }
    }
}

// Occupancy is what the row gathers live on (bytes in flight towards L2): measured at Reddit shape, dim 16:
// 4 CTAs/SM (61 registers) 700 us, 5 CTAs/SM (48 registers) 595 us per launch.
#ifndef GCNK_GATHER_MIN_CTAS
#define GCNK_GATHER_MIN_CTAS 8
#endif
#ifndef GCNK_GATHER_DEFAULT_VARIANT
#define GCNK_GATHER_DEFAULT_VARIANT 2
#endif
// IDX4: 0 = coalesced index load + shuffles; 1 = int4 index reads, 32 registers (8 CTAs/SM); 2 = int4 index reads with
// the four row reads of a chunk in flight together, 40 registers (6 CTAs/SM); 3 = two chunks (eight row reads) in
// flight, 64 registers (4 CTAs/SM).
template <int VEC, int LPR, int NACC, bool EXACT, int IDX4>
__global__ void __launch_bounds__(THREADS, NACC != 1 ? 1 : IDX4 == 3 ? 4 : IDX4 == 2 ? 6 : GCNK_GATHER_MIN_CTAS) gather_kernel(const GatherArgs a) {
    extern __shared__ float smem[];   // heavy rows only: [WARPS][dim]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Acc<VEC> acc[NACC];
    wait_for_peers(a);

    if ((int)blockIdx.x < a.n_heavy) {
        // a whole CTA on one high-degree row: warps take interleaved 32-edge chunks, partials are
        // combined through shared memory in warp order (fixed => deterministic)
        const int s = a.heavy_rows[blockIdx.x];
        const int beg = a.indptr[s], end = a.indptr[s + 1];
#pragma unroll
        for (int t = 0; t < NACC; t++) {
            acc[t].zero();
            const int u = (lane % LPR + LPR * t) * VEC;
            if (a.init && warp == 0 && lane < LPR && u < a.dim) acc[t].load(a.init + (size_t)s * a.dim + u);   // the partial sum so far
        }
        if constexpr (IDX4 != 0) accumulate_idx4<EXACT, WARPS, IDX4 == 3 ? 8 : IDX4 == 2 ? 4 : 0>(acc[0], a, beg, end, warp, lane);
        else accumulate<VEC, LPR, NACC, EXACT>(acc, a, beg, end, warp, WARPS, lane);
        reduce_groups<VEC, LPR, NACC>(acc);
        const int q = lane % LPR;
        if (lane < LPR) {
#pragma unroll
            for (int t = 0; t < NACC; t++) {
                const int u = (q + LPR * t) * VEC;
#pragma unroll
                for (int v = 0; v < VEC; v++)
                    if (u + v < a.dim) smem[warp * a.dim + u + v] = acc[t].get(v);
            }
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int t = 0; t < NACC; t++) {
                const int u = (q + LPR * t) * VEC;
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    float sum = 0.f;
                    if (lane < LPR && u + v < a.dim)
                        for (int w = 0; w < WARPS; w++) sum += smem[w * a.dim + u + v];
                    acc[t].set(v, sum);
                }
            }
            epilogue<VEC, LPR, NACC>(acc, a, s, lane);
        }
        return;
    }

    const int bin = ((int)blockIdx.x - a.n_heavy) * WARPS + warp;
    const int r_end = a.bin_ptr[bin + 1];
    for (int r = a.bin_ptr[bin]; r < r_end; r++) {
        const int s = a.bin_rows[r];
        const int beg = a.indptr[s], end = a.indptr[s + 1];
#pragma unroll
        for (int t = 0; t < NACC; t++) {
            acc[t].zero();
            const int u = (lane % LPR + LPR * t) * VEC;
            if (a.init && lane < LPR && u < a.dim) acc[t].load(a.init + (size_t)s * a.dim + u);                  // the partial sum so far
        }
        if constexpr (IDX4 != 0) accumulate_idx4<EXACT, 1, IDX4 == 3 ? 8 : IDX4 == 2 ? 4 : 0>(acc[0], a, beg, end, 0, lane);
        else accumulate<VEC, LPR, NACC, EXACT>(acc, a, beg, end, 0, 1, lane);
        reduce_groups<VEC, LPR, NACC>(acc);
        epilogue<VEC, LPR, NACC>(acc, a, s, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// GraphSum with the exchange of its source fused in (row-partitioned runs, width 12 / 16): ONE launch pushes this rank's
// rows of the gather source to the peers over NVLink and aggregates, overlapping the two.
//   CTAs [0, x_ctas)   copy the rank's finished rows (x_vec float4) into every peer's buffer with coalesced 16-byte stores;
//                      the last of them to finish publishes x_value in the peers' flag slots.
//   the other CTAs     aggregate as gather_kernel does, on the ROTATED row order (own columns, higher ranks, lower ranks),
//                      in one pass per row: own columns need nothing from anybody; before the first 32-entry chunk that
//                      reaches the higher ranks' columns the warp waits for their flags, likewise for the lower ranks —
//                      by which time the pushes, which started together with this kernel on every rank, have landed.
//                      The chunking does not depend on when the peers arrive: results stay reproducible.
// Nothing of a peer's block is read before its flag has been seen (acquire at system scope); partition cuts are even, so
// no 128-byte line of a 64-byte-row source holds rows of two ranks.
// Polls with plain volatile loads (served by L2, where the peer's flag store lands): tens of thousands of warps pass through
// here, and an acquire at system scope per warp — a full fence each — was measured to cost more than the exchange it
// guards.  Ordering still holds: the writer fences (system scope) between its rows and its flag; this side reads a peer's
// rows only after the loop has seen the flag (a control dependence the hardware does not speculate past), and has never
// touched those lines before, so they cannot be stale in L1.
__device__ __forceinline__ void x_wait(const GatherArgs &a, int lo, int hi, int lane) {
    const int r = lo + lane;
    if (r < hi) {
        const volatile int *f = a.wait_flags + r;
        if (*f < a.wait_value) {
            const long long t0 = clock64();
            while (*f < a.wait_value) {
                if (clock64() - t0 > a.wait_limit) { *a.wait_err = 1; break; }
                __nanosleep(32);
            }
        }
    }
    __syncwarp();
}

// The BATCH == 4 loop of accumulate_idx4 over the ROTATED row [beg, end) = [own columns | m1: higher ranks | m2: lower
// ranks], in ONE pass: before a 32-entry chunk that reaches into a part whose owners have not been seen yet, the warp
// waits for their flags (`state`: 0 nothing seen, 1 higher ranks seen, 2 all seen — kept across rows).  The chunking, hence
// the order of the additions, does not depend on `state`: results are the same whenever the peers arrive.
template <bool EXACT, int STEP>
__device__ __forceinline__ void accumulate_x(Acc<4> &acc, const GatherArgs &a, int beg, int end, int m1, int m2, int first, int lane, int &state) {
    constexpr int STRIDE = 32 * STEP;
    const int g4 = (lane >> 2) * 4, q = lane & 3;
    const int dim = EXACT ? 16 : a.dim;
    const int base = (beg & ~3) + 32 * first;
    int left = end - base;
    if (left <= 0) return;
    const int *ip = a.indices + base + g4;
    const float *in_q = a.in + q * 4;
    const bool active = EXACT || q * 4 < dim;              // (lanes beyond the row width still take part in the waits)
    int4 idx = make_int4(0, 0, 0, 0);
    if (left > g4) idx = load_idx4(ip);
    int4 idx_next = idx;
    if (left - STRIDE > g4) idx_next = load_idx4(ip + STRIDE);
    auto need = [&](int chunk_end) {
        if (state < 1 && chunk_end > m1) { x_wait(a, a.x_rank + 1, a.x_world, lane); state = 1; }
        if (state < 2 && chunk_end > m2) { x_wait(a, 0, a.x_rank, lane); state = 2; }
    };
    if (state < 2) need(min(end, base + 32));
    if (active) {
        if (base >= beg && left >= 32) gather4_full<true>(acc, in_q, dim, idx);
        else gather4_edge(acc, in_q, dim, idx, beg - base - g4, left - g4);
    }
#pragma unroll 1
    while (left > STRIDE) {
        ip += STRIDE;
        left -= STRIDE;
        idx = idx_next;
        if (left - STRIDE > g4) idx_next = load_idx4(ip + STRIDE);
        if (state < 2) need(min(end, end - left + 32));
        if (active) {
            if (left >= 32) gather4_full<true>(acc, in_q, dim, idx);
            else gather4_edge(acc, in_q, dim, idx, 0, left - g4);
        }
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(THREADS, 6) xgather_kernel(const GatherArgs a) {
    extern __shared__ float smem[];   // heavy rows only: [WARPS][dim]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x < a.x_ctas) {
        // ---------------- pushers
        const size_t stride = (size_t)a.x_ctas * THREADS;
        for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < a.x_vec; i += stride) {
            const float4 v = a.x_src[i];
#pragma unroll 1
            for (int p = 0; p < a.x_peers; p++) a.x_dst[p][i] = v;
        }
        __shared__ bool s_last;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned prev = atomicAdd(a.x_counter, 1u);
            s_last = prev == (unsigned)a.x_ctas - 1;
            if (s_last) *a.x_counter = 0;
        }
        __syncthreads();
        if (s_last && (int)threadIdx.x < a.x_peers) {
            __threadfence_system();
            *reinterpret_cast<volatile int *>(a.x_flag[threadIdx.x]) = a.x_value;
        }
        return;
    }
    const int bid = (int)blockIdx.x - a.x_ctas;
    Acc<4> acc[1];
    if (bid < a.n_heavy) {
        const int s = a.heavy_rows[bid];
        const int beg = a.indptr[s], end = a.indptr[s + 1];
        int state = 0;
        acc[0].zero();
        accumulate_x<EXACT, WARPS>(acc[0], a, beg, end, a.seg1[s], a.seg2[s], warp, lane, state);
        reduce_groups<4, 4, 1>(acc);
        const int q = lane % 4;
        if (lane < 4) {
#pragma unroll
            for (int v = 0; v < 4; v++)
                if (q * 4 + v < a.dim) smem[warp * a.dim + q * 4 + v] = acc[0].get(v);
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int v = 0; v < 4; v++) {
                float sum = 0.f;
                if (lane < 4 && q * 4 + v < a.dim)
                    for (int w = 0; w < WARPS; w++) sum += smem[w * a.dim + q * 4 + v];
                acc[0].set(v, sum);
            }
            epilogue<4, 4, 1>(acc, a, s, lane);
        }
        return;
    }
    const int bin = (bid - a.n_heavy) * WARPS + warp;
    const int r_end = a.bin_ptr[bin + 1];
    int state = 0;
    for (int r = a.bin_ptr[bin]; r < r_end; r++) {
        const int s = a.bin_rows[r];
        const int beg = a.indptr[s], end = a.indptr[s + 1];
        acc[0].zero();
        if (state < 2) accumulate_x<EXACT, 1>(acc[0], a, beg, end, a.seg1[s], a.seg2[s], 0, lane, state);
        else accumulate_idx4<EXACT, 1, 4>(acc[0], a, beg, end, 0, lane);       // everything has arrived: the plain loop (same order)
        reduce_groups<4, 4, 1>(acc);
        epilogue<4, 4, 1>(acc, a, s, lane);
    }
}

// Which index-fetch variant the dim 12 / 16 gather uses (see gather_kernel): GCNK_GATHER_IDX4 in the environment, or
// gcnk_gather_variant() at run time.
int g_gather_variant = -1;
int gather_variant() {
    if (g_gather_variant < 0) {
        const char *e = getenv("GCNK_GATHER_IDX4");
        g_gather_variant = (e && *e >= '0' && *e <= '3') ? *e - '0' : GCNK_GATHER_DEFAULT_VARIANT;
    }
    return g_gather_variant;
}

// True iff [p, p + bytes) lies inside one device allocation (cuMemGetAddressRange through the runtime's driver entry
// point, so libcuda is not a link dependency).
bool device_readable(const void *p, size_t bytes) {
    typedef CUresult (*range_fn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static range_fn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) f = nullptr;
        cudaGetLastError();
        return reinterpret_cast<range_fn>(f);
    }();
    if (!fn || !p) return false;
    CUdeviceptr base = 0;
    size_t size = 0;
    if (fn(&base, &size, (CUdeviceptr)(uintptr_t)p) != CUDA_SUCCESS) return false;
    return (uintptr_t)p + bytes <= (uintptr_t)base + size;
}
int idx4_readable(const int *indices, int64_t entries) {
    return reinterpret_cast<uintptr_t>(indices) % 16 == 0 && device_readable(indices, sizeof(int) * (size_t)((entries + 3) / 4 * 4));
}

template <int VEC, int LPR, int NACC>
int launch_variant(const gcnk_graph *g, const GatherArgs &a, cudaStream_t st) {
    const int grid = g->n_heavy + g->n_bins / WARPS;
    if (grid == 0) return GCNK_OK;
    const size_t smem = g->n_heavy ? sizeof(float) * WARPS * (size_t)a.dim : 0;
    if constexpr (LPR == 4 && VEC == 4 && NACC == 1) {
        const int variant = g->idx4_ok ? gather_variant() : 0;
        const bool exact = a.dim == VEC * LPR * NACC;
        if (variant == 1) {
            if (exact) { prefer_carveout(gather_kernel<VEC, LPR, NACC, true, 1>); gather_kernel<VEC, LPR, NACC, true, 1><<<grid, THREADS, smem, st>>>(a); }
            else { prefer_carveout(gather_kernel<VEC, LPR, NACC, false, 1>); gather_kernel<VEC, LPR, NACC, false, 1><<<grid, THREADS, smem, st>>>(a); }
        } else if (variant == 2) {
            if (exact) { prefer_carveout(gather_kernel<VEC, LPR, NACC, true, 2>); gather_kernel<VEC, LPR, NACC, true, 2><<<grid, THREADS, smem, st>>>(a); }
            else { prefer_carveout(gather_kernel<VEC, LPR, NACC, false, 2>); gather_kernel<VEC, LPR, NACC, false, 2><<<grid, THREADS, smem, st>>>(a); }
        } else if (variant == 3) {
            if (exact) { prefer_carveout(gather_kernel<VEC, LPR, NACC, true, 3>); gather_kernel<VEC, LPR, NACC, true, 3><<<grid, THREADS, smem, st>>>(a); }
            else { prefer_carveout(gather_kernel<VEC, LPR, NACC, false, 3>); gather_kernel<VEC, LPR, NACC, false, 3><<<grid, THREADS, smem, st>>>(a); }
        }
        if (variant) { GCNK_LAUNCHED(); return GCNK_OK; }
    }
    if (a.dim == VEC * LPR * NACC) { prefer_carveout(gather_kernel<VEC, LPR, NACC, true, 0>); gather_kernel<VEC, LPR, NACC, true, 0><<<grid, THREADS, smem, st>>>(a); }
    else { prefer_carveout(gather_kernel<VEC, LPR, NACC, false, 0>); gather_kernel<VEC, LPR, NACC, false, 0><<<grid, THREADS, smem, st>>>(a); }
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int launch_gather(const gcnk_graph *g, GatherArgs a, cudaStream_t st) {
    const int dim = a.dim;
    a.indptr = g->indptr; a.indices = g->indices; a.dinv = g->dinv;
    a.heavy_rows = g->heavy_rows; a.n_heavy = g->n_heavy; a.bin_ptr = g->bin_ptr; a.bin_rows = g->bin_rows;
    a.mask_stride = mask_stride_bits(dim);
    a.init = t_init;
    t_init = nullptr;
    if (t_xchg.armed) {
        t_xchg.armed = false;
        if (!(g->seg1 && g->seg2 && g->idx4_ok && (dim == 16 || dim == 12) && reinterpret_cast<uintptr_t>(a.in) % 16 == 0 &&
              reinterpret_cast<uintptr_t>(a.out) % 16 == 0 && t_xchg.n_floats % 4 == 0 && !a.init)) {
            set_error("gather: the fused exchange needs a rotated graph (gcnk_graph_rotate), width 12 or 16 and 16-byte-aligned buffers");
            return GCNK_EUNSUPPORTED;
        }
        a.x_src = reinterpret_cast<const float4 *>(t_xchg.src); a.x_vec = t_xchg.n_floats / 4; a.x_peers = t_xchg.peers;
        for (int i = 0; i < t_xchg.peers; i++) { a.x_dst[i] = reinterpret_cast<float4 *>(t_xchg.dst[i]); a.x_flag[i] = t_xchg.flag[i]; }
        a.x_counter = t_xchg.counter; a.x_value = t_xchg.value; a.x_rank = t_xchg.rank; a.x_world = t_xchg.world;
        a.x_ctas = (int)std::max<size_t>(1, std::min<size_t>((a.x_vec + 2 * THREADS - 1) / (2 * THREADS), (size_t)sm_count()));   // short-lived: they leave their slots to the gather CTAs
        a.wait_flags = t_xchg.wait_flags; a.wait_value = t_xchg.value; a.wait_err = t_xchg.err; a.wait_limit = peer_spin_cycles();
        a.seg1 = g->seg1; a.seg2 = g->seg2;
        const int grid = a.x_ctas + g->n_heavy + g->n_bins / WARPS;
        const size_t smem = g->n_heavy ? sizeof(float) * WARPS * (size_t)dim : 0;
        if (dim == 16) xgather_kernel<true><<<grid, THREADS, smem, st>>>(a);
        else xgather_kernel<false><<<grid, THREADS, smem, st>>>(a);
        GCNK_LAUNCHED();
        return GCNK_OK;
    }
    if (t_wait.flags) {
        a.wait_flags = t_wait.flags; a.wait_n = t_wait.n; a.wait_skip = t_wait.skip; a.wait_value = t_wait.value;
        a.wait_err = t_wait.err; a.wait_limit = peer_spin_cycles();
        t_wait.flags = nullptr;
    }
    // No mirrored epilogue here on purpose: the 7 extra pointers cost the gather 19 registers (61 -> 80, one CTA per
    // SM less, +25 % run time measured), and posted 64-byte remote stores back up the load/store unit the row gathers
    // depend on.  A pending gcnk_mirror_next registration is left for the caller, which pushes the finished rows to
    // the peers with one coalesced copy kernel (gcnk_peer_push).
    const bool vec4 = dim % 4 == 0 && (reinterpret_cast<uintptr_t>(a.in) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(a.out) % 16 == 0);
    if (vec4) {
        const int units = dim / 4;
        if (units <= 1) return launch_variant<4, 1, 1>(g, a, st);
        if (units <= 2) return launch_variant<4, 2, 1>(g, a, st);
        if (units <= 4) return launch_variant<4, 4, 1>(g, a, st);
        if (units <= 8) return launch_variant<4, 8, 1>(g, a, st);
        if (units <= 16) return launch_variant<4, 16, 1>(g, a, st);
        if (units <= 32) return launch_variant<4, 32, 1>(g, a, st);
        if (units <= 64) return launch_variant<4, 32, 2>(g, a, st);
        if (units <= 128) return launch_variant<4, 32, 4>(g, a, st);
        if (units <= 256) return launch_variant<4, 32, 8>(g, a, st);
    } else {
        if (dim <= 1) return launch_variant<1, 1, 1>(g, a, st);
        if (dim <= 2) return launch_variant<1, 2, 1>(g, a, st);
        if (dim <= 4) return launch_variant<1, 4, 1>(g, a, st);
        if (dim <= 8) return launch_variant<1, 8, 1>(g, a, st);
        if (dim <= 16) return launch_variant<1, 16, 1>(g, a, st);
        if (dim <= 32) return launch_variant<1, 32, 1>(g, a, st);
        if (dim <= 64) return launch_variant<1, 32, 2>(g, a, st);
        if (dim <= 128) return launch_variant<1, 32, 4>(g, a, st);
        if (dim <= 256) return launch_variant<1, 32, 8>(g, a, st);
        if (dim <= 512) return launch_variant<1, 32, 16>(g, a, st);
        if (dim <= 1024) return launch_variant<1, 32, 32>(g, a, st);
    }
    set_error("gather: dim %d unsupported (max 1024, as the reference's GPU path, cuda_module.cu:79)", dim);
    return GCNK_EUNSUPPORTED;
}

__global__ void dinv_kernel(const int *indptr, float *dinv, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int deg = indptr[i + 1] - indptr[i];
        // 1/sqrt(deg) in IEEE fp32; a zero-degree row (impossible after the parser's self loop) yields 0
        dinv[i] = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.f;
    }
}

// symmetric iff every stored (s,d) has a stored (d,s); rows must be sorted after the leading self
// loop for the binary search to succeed — an unsorted input simply reports "not verified".
__global__ void symmetry_kernel(const int *indptr, const int *indices, int n, int n_cols, int *asym) {
    const int s = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (s >= n) return;
    const int row_beg = indptr[s];
    for (int e = row_beg + lane; e < indptr[s + 1]; e += 32) {
        const int d = indices[e];
        if (d == s) continue;
        if (d < 0 || d >= n_cols || d >= n) { atomicOr(asym, 1); continue; }
        // A repeated neighbour (the parser keeps duplicates, as the reference's does, parser.cpp:31-40) makes
        // A_hat[s][d] = k / sqrt(deg_s deg_d): symmetric only if row d repeats s as often.  Multiplicities are not
        // compared here — any repeated entry reports "not verified", and such inputs run the modules plan.
        if (e > row_beg + 1 && indices[e - 1] == d) { atomicOr(asym, 1); continue; }
        int lo = indptr[d], hi = indptr[d + 1];
        if (lo < hi && indices[lo] == d) lo++;             // skip the leading self loop
        bool found = false;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1, v = indices[mid];
            if (v == s) { found = true; break; }
            if (v < s) lo = mid + 1; else hi = mid;
        }
        if (!found) atomicOr(asym, 1);
    }
}

__global__ void scale_rows_kernel(const float *__restrict__ dinv, const float *__restrict__ in, float *__restrict__ out,
                                  int64_t total, int dim) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) out[i] = dinv[i / dim] * in[i];
}

// gcnk_graphsum at a width that is not a multiple of 4 (41 and 47 are the class widths of the benchmark shapes): rows padded to
// a 16-byte pitch, so that the gather moves them with 128-bit loads, several edges per instruction, instead of one edge per
// instruction with scalar loads (4.4 ms -> see DESIGN 4 at width 41 on the Reddit-shape graph).  The pre-scale pass writes a copy anyway.
__global__ void scale_rows_pad_kernel(const float *__restrict__ dinv, const float *__restrict__ in, float *__restrict__ out,
                                      int64_t total, int dim, int pitch) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int64_t r = i / pitch;
        const int c = (int)(i - r * pitch);
        out[i] = c < dim ? dinv[r] * in[r * dim + c] : 0.f;
    }
}
__global__ void unpad_rows_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t total, int dim, int pitch) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int64_t r = i / dim;
        out[i] = in[r * pitch + (i - r * dim)];
    }
}

// Rotated order of a row-partition's CSR slice (gcnk_graph_rotate): a warp per row rewrites its entries as
// [columns in [lo, hi) | columns >= hi | columns < lo], each class in its original order, and records the two boundaries.
__global__ void rotate_rows_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, int n, int lo, int hi, int *__restrict__ out,
                                   int *__restrict__ seg1, int *__restrict__ seg2) {
    const int s = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (s >= n) return;
    const int beg = indptr[s], end = indptr[s + 1];
    int w = beg;
    for (int cls = 0; cls < 3; cls++) {
        for (int e0 = beg; e0 < end; e0 += 32) {
            const int e = e0 + lane;
            const int d = e < end ? indices[e] : -1;
            const int c = d < 0 ? -1 : (d >= lo && d < hi) ? 0 : d >= hi ? 1 : 2;
            const unsigned m = __ballot_sync(FULL, c == cls);
            if (c == cls) out[w + __popc(m & ((1u << lane) - 1u))] = d;
            w += __popc(m);
        }
        if (lane == 0) { if (cls == 0) seg1[s] = w; else if (cls == 1) seg2[s] = w; }
    }
}
// boundaries of an already rotated row (a column-filtered view of a rotated graph keeps the order)
__global__ void segment_rows_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, int n, int lo, int hi, int *__restrict__ seg1,
                                    int *__restrict__ seg2) {
    const int s = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (s >= n) return;
    int own = 0, up = 0;
    for (int e = indptr[s] + lane; e < indptr[s + 1]; e += 32) {
        const int d = indices[e];
        own += d >= lo && d < hi;
        up += d >= hi;
    }
    own = warp_sum_int(own); up = warp_sum_int(up);
    if (lane == 0) { seg1[s] = indptr[s] + own; seg2[s] = indptr[s] + own + up; }
}

// Column-filtered views (gcnk_graph_create_view): a warp per row counts / compacts the entries whose column flag is set,
// in order (ballot + prefix popcount)
__global__ void view_count_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const int *__restrict__ col_keep, int n, int n_cols,
                                  int *__restrict__ cnt) {
    const int s = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (s >= n) return;
    int c = 0;
    for (int e = indptr[s] + lane; e < indptr[s + 1]; e += 32) {
        const int d = indices[e];
        c += d >= 0 && d < n_cols && col_keep[d] != 0;
    }
    c = warp_sum_int(c);
    if (lane == 0) cnt[s] = c;
}
__global__ void view_fill_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, const int *__restrict__ col_keep,
                                 const int *__restrict__ new_ptr, int n, int n_cols, int *__restrict__ out) {
    const int s = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (s >= n) return;
    int w = new_ptr[s];
    const int beg = indptr[s], end = indptr[s + 1];
    for (int e0 = beg; e0 < end; e0 += 32) {
        const int e = e0 + lane;
        const int d = e < end ? indices[e] : -1;
        const bool keep = d >= 0 && d < n_cols && col_keep[d] != 0;
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) out[w + __popc(m & ((1u << lane) - 1u))] = d;
        w += __popc(m);
    }
}

// Static schedule over the given rows: heavy rows -> one CTA each; the rest -> LPT bins, one warp per bin.
int build_schedule(gcnk_graph *g, const std::vector<int> &indptr, const std::vector<int> &rows, cudaStream_t st) {
    // A warp walks its row 32 entries at a time and every step is an L2 round trip, so the longest single-warp row
    // is a latency floor (2,048 entries ~ 45 us).  That is invisible when the launch has hundreds of microseconds of
    // throughput-bound work, but a rank of an 8-way partition has ~100 us in total: the CTA-per-row threshold
    // therefore shrinks with the work of the launch (8 warps then share the row's 32-entry steps).
    int64_t work = 0;
    for (int i : rows) work += indptr[i + 1] - indptr[i];
    const int heavy_degree = (int)std::max<int64_t>(256, std::min<int64_t>(HEAVY_DEGREE, work / ((int64_t)sm_count() * 128)));
    std::vector<int> heavy, light;
    int max_deg = 0;
    for (int i : rows) {
        const int deg = indptr[i + 1] - indptr[i];
        max_deg = std::max(max_deg, deg);
        (deg > heavy_degree ? heavy : light).push_back(i);
    }
    g->max_degree = max_deg;
    g->n_rows_scheduled = (int)rows.size();
    auto deg_of = [&](int r) { return indptr[r + 1] - indptr[r]; };
    std::stable_sort(heavy.begin(), heavy.end(), [&](int x, int y) { return deg_of(x) > deg_of(y); });
    std::stable_sort(light.begin(), light.end(), [&](int x, int y) { return deg_of(x) > deg_of(y); });
    // Bins are equal-sized by construction, so a launch runs in waves of (resident warps per SM x SMs) bins: the bin
    // count is a whole number of waves of the dim-16 kernel (4-5 bins per resident warp keeps the tail short without
    // shrinking bins below a few rows).  GCNK_GATHER_BINS_PER_SM overrides it for experiments.
    int bins_per_sm = gather_variant() == 2 ? 6 * WARPS * 5 : 8 * WARPS * 4;   // (variant 3: 4 CTAs/SM, 8 waves)
    if (const char *e = getenv("GCNK_GATHER_BINS_PER_SM")) { const int v = atoi(e); if (v >= WARPS && v <= 4096) bins_per_sm = v; }
    int n_bins = sm_count() * bins_per_sm;
    n_bins = std::min<int64_t>(n_bins, std::max<int64_t>((int64_t)light.size(), 1));
    n_bins = (n_bins + WARPS - 1) / WARPS * WARPS;
    if (light.empty()) n_bins = 0;
    std::vector<std::vector<int>> bins(n_bins);
    {
        typedef std::pair<int64_t, int> Item;   // (load, bin) — min-heap on load, ties by bin id => deterministic
        std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
        for (int b = 0; b < n_bins; b++) heap.push({0, b});
        for (int r : light) {
            Item it = heap.top(); heap.pop();
            bins[it.second].push_back(r);
            it.first += deg_of(r) + ROW_OVERHEAD;
            heap.push(it);
        }
    }
    std::vector<int> bin_ptr(n_bins + 1, 0), bin_rows;
    bin_rows.reserve(light.size());
    for (int b = 0; b < n_bins; b++) {
        bin_rows.insert(bin_rows.end(), bins[b].begin(), bins[b].end());
        bin_ptr[b + 1] = (int)bin_rows.size();
    }
    g->n_heavy = (int)heavy.size(); g->n_bins = n_bins;
    GCNK_CUDA(cudaMalloc(&g->heavy_rows, sizeof(int) * std::max<size_t>(heavy.size(), 1)));
    GCNK_CUDA(cudaMalloc(&g->bin_ptr, sizeof(int) * (n_bins + 1)));
    GCNK_CUDA(cudaMalloc(&g->bin_rows, sizeof(int) * std::max<size_t>(bin_rows.size(), 1)));
    if (!heavy.empty()) GCNK_CUDA(cudaMemcpyAsync(g->heavy_rows, heavy.data(), sizeof(int) * heavy.size(), cudaMemcpyHostToDevice, st));
    GCNK_CUDA(cudaMemcpyAsync(g->bin_ptr, bin_ptr.data(), sizeof(int) * bin_ptr.size(), cudaMemcpyHostToDevice, st));
    if (!bin_rows.empty()) GCNK_CUDA(cudaMemcpyAsync(g->bin_rows, bin_rows.data(), sizeof(int) * bin_rows.size(), cudaMemcpyHostToDevice, st));
    GCNK_CUDA(cudaStreamSynchronize(st));
    return GCNK_OK;
}

}  // namespace

extern "C" {

int gcnk_graph_create(gcnk_graph **out, const int *d_indptr, const int *d_indices, int n, int64_t nnz, int n_cols,
                      const float *d_dinv_global, gcnk_stream_t stream) {
    GCNK_REQUIRE(out && d_indptr && (d_indices || nnz == 0) && n >= 0 && n_cols >= n, "bad arguments");
    GCNK_REQUIRE(n_cols == n || d_dinv_global, "a row partition (n_cols > n) needs the global d^-1/2 vector");
    cudaStream_t st = S(stream);
    gcnk_graph *g = new gcnk_graph;
    g->indptr = d_indptr; g->indices = d_indices; g->n = n; g->n_cols = n_cols; g->nnz = nnz;

    std::vector<int> indptr((size_t)n + 1, 0);
    GCNK_CUDA(cudaMemcpyAsync(indptr.data(), d_indptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
    GCNK_CUDA(cudaMalloc(&g->dinv, sizeof(float) * std::max(n, 1)));
    if (n) { dinv_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_indptr, g->dinv, n); GCNK_LAUNCHED(); }
    g->dinv_cols = d_dinv_global ? d_dinv_global : g->dinv;
    int *d_asym = nullptr, asym = 0;
    GCNK_CUDA(cudaMalloc(&d_asym, sizeof(int)));
    GCNK_CUDA(cudaMemsetAsync(d_asym, 0, sizeof(int), st));
    if (n && n_cols == n) { symmetry_kernel<<<(n + 7) / 8, 256, 0, st>>>(d_indptr, d_indices, n, n_cols, d_asym); GCNK_LAUNCHED(); }
    GCNK_CUDA(cudaMemcpyAsync(&asym, d_asym, sizeof(int), cudaMemcpyDeviceToHost, st));
    GCNK_CUDA(cudaStreamSynchronize(st));
    GCNK_CUDA(cudaFree(d_asym));
    g->symmetric = (n_cols == n) && !asym;
    g->idx4_ok = nnz > 0 && idx4_readable(d_indices, nnz);
    if (n && indptr[n] != nnz) { delete g; set_error("gcnk_graph_create: indptr[n]=%d != nnz=%lld", indptr[n], (long long)nnz); return GCNK_EINVAL; }

    std::vector<int> rows((size_t)n);
    for (int i = 0; i < n; i++) rows[i] = i;
    {
        const int rc = build_schedule(g, indptr, rows, st);
        if (rc) { delete g; return rc; }
    }
    *out = g;
    return GCNK_OK;
}


int gcnk_graph_create_view(gcnk_graph **out, const gcnk_graph *base, const int *d_row_keep, const int *d_col_keep,
                           gcnk_stream_t stream) {
    GCNK_REQUIRE(out && base, "needs a base graph");
    GCNK_REQUIRE(!base->base || !d_col_keep, "a view of a view can only restrict the rows (it shares the parent's filtered CSR)");
    cudaStream_t st = S(stream);
    const int n = base->n;
    gcnk_graph *g = new gcnk_graph;
    *g = *base;
    g->base = base->base ? base->base : base;        // d^-1/2 always comes from the root graph
    g->heavy_rows = nullptr; g->bin_ptr = nullptr; g->bin_rows = nullptr; g->scratch = nullptr; g->scratch_elems = 0;
    g->own_indptr = nullptr; g->own_indices = nullptr;   // (a row view of a column view borrows that view's CSR: it must outlive this one)
    g->seg_owned = false;                                 // ... and the parent's rotation boundaries

    std::vector<int> indptr((size_t)n + 1, 0), row_keep;
    if (d_row_keep) {
        row_keep.resize((size_t)n);
        GCNK_CUDA(cudaMemcpyAsync(row_keep.data(), d_row_keep, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    if (d_col_keep) {
        // Drop the entries whose column is filtered out; row order and the order inside a row are preserved.  Counted and
        // compacted on the device (a warp per row); only the n + 1 row offsets travel to the host, for the prefix sum and
        // the schedule — not the index array (459 MB at Reddit shape).
        int *d_cnt = nullptr;
        GCNK_CUDA(cudaMalloc(&d_cnt, sizeof(int) * ((size_t)n + 1)));
        GCNK_CUDA(cudaMalloc(&g->own_indptr, sizeof(int) * ((size_t)n + 1)));
        if (n) { view_count_kernel<<<(n + 7) / 8, 256, 0, st>>>(base->indptr, base->indices, d_col_keep, n, base->n_cols, d_cnt); GCNK_LAUNCHED(); }
        std::vector<int> cnt((size_t)n + 1, 0);
        GCNK_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
        GCNK_CUDA(cudaStreamSynchronize(st));
        int64_t w = 0;
        for (int i = 0; i < n; i++) { indptr[i] = (int)w; w += cnt[i]; }
        indptr[n] = (int)w;
        GCNK_CUDA(cudaMalloc(&g->own_indices, sizeof(int) * ((size_t)w + 4)));     // + 4: the int4 index reads round up
        GCNK_CUDA(cudaMemsetAsync(g->own_indices + w, 0, sizeof(int) * 4, st));
        GCNK_CUDA(cudaMemcpyAsync(g->own_indptr, indptr.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
        if (n) { view_fill_kernel<<<(n + 7) / 8, 256, 0, st>>>(base->indptr, base->indices, d_col_keep, g->own_indptr, n, base->n_cols, g->own_indices); GCNK_LAUNCHED(); }
        GCNK_CUDA(cudaStreamSynchronize(st));
        GCNK_CUDA(cudaFree(d_cnt));
        g->indptr = g->own_indptr; g->indices = g->own_indices; g->nnz = w;
        g->idx4_ok = w > 0 && idx4_readable(g->own_indices, w);
        g->symmetric = 0;
        if (base->seg1) {                     // a view of a rotated graph: its own boundaries (the filter keeps the order)
            g->seg1 = nullptr; g->seg2 = nullptr;
            GCNK_CUDA(cudaMalloc(&g->seg1, sizeof(int) * std::max(n, 1)));
            GCNK_CUDA(cudaMalloc(&g->seg2, sizeof(int) * std::max(n, 1)));
            g->seg_owned = true;
            if (n) { segment_rows_kernel<<<(n + 7) / 8, 256, 0, st>>>(g->own_indptr, g->own_indices, n, base->rot_lo, base->rot_hi, g->seg1, g->seg2); GCNK_LAUNCHED(); }
            GCNK_CUDA(cudaStreamSynchronize(st));
        }
    } else {
        GCNK_CUDA(cudaMemcpyAsync(indptr.data(), base->indptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
        GCNK_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<int> rows;
    rows.reserve((size_t)n);
    for (int i = 0; i < n; i++)
        if (!d_row_keep || row_keep[i]) rows.push_back(i);
    if (d_row_keep) {
        int64_t nnz_rows = 0;
        for (int i : rows) nnz_rows += indptr[i + 1] - indptr[i];
        g->nnz = nnz_rows;                    // entries this view's launches actually touch
    }
    const int rc = build_schedule(g, indptr, rows, st);
    if (rc) { gcnk_graph_destroy(g); return rc; }
    *out = g;
    return GCNK_OK;
}

int gcnk_graph_destroy(gcnk_graph *g) {
    if (!g) return GCNK_OK;
    if (!g->base) cudaFree(g->dinv);                  // views borrow d^-1/2 from the root graph
    cudaFree(g->heavy_rows); cudaFree(g->bin_ptr); cudaFree(g->bin_rows); cudaFree(g->scratch);
    cudaFree(g->own_indptr); cudaFree(g->own_indices);
    if (g->seg_owned) { cudaFree(g->seg1); cudaFree(g->seg2); }
    delete g;
    return GCNK_OK;
}

int gcnk_graph_dinv(const gcnk_graph *g, const float **d) { GCNK_REQUIRE(g && d, "null"); *d = g->dinv; return GCNK_OK; }

int gcnk_graph_stats(const gcnk_graph *g, int *n, int64_t *nnz, int *max_degree, int *is_symmetric, int *n_bins) {
    GCNK_REQUIRE(g, "null graph");
    if (n) *n = g->n;
    if (nnz) *nnz = g->nnz;
    if (max_degree) *max_degree = g->max_degree;
    if (is_symmetric) *is_symmetric = g->symmetric;
    if (n_bins) *n_bins = g->n_bins;
    return GCNK_OK;
}

int gcnk_scale_rows(const float *dinv, const float *in, float *out, int rows, int dim, gcnk_stream_t stream) {
    GCNK_REQUIRE(dinv && in && out && rows >= 0 && dim > 0, "bad arguments");
    const int64_t total = (int64_t)rows * dim;
    if (!total) return GCNK_OK;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    scale_rows_kernel<<<grid, 256, 0, S(stream)>>>(dinv, in, out, total, dim);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_gather_plain(const gcnk_graph *g, const float *in_scaled, float *out, int dim, gcnk_stream_t stream) {
    GCNK_REQUIRE(g && in_scaled && out && dim > 0, "bad arguments");
    GatherArgs a = {};
    a.in = in_scaled; a.out = out; a.dim = dim; a.mode = MODE_PLAIN; a.scale = 1.f;
    return launch_gather(g, a, S(stream));
}

int gcnk_gather_relu_drop(const gcnk_graph *g, const float *in_scaled, float *out_scaled, const uint32_t *drop_bits,
                          uint32_t *mask_bits, float scale, int dim, gcnk_stream_t stream) {
    GCNK_REQUIRE(g && in_scaled && out_scaled && dim > 0, "bad arguments");
    GatherArgs a = {};
    a.in = in_scaled; a.out = out_scaled; a.dim = dim; a.mode = MODE_RELU_DROP;
    a.drop_bits = drop_bits; a.mask_out = mask_bits; a.scale = scale;
    return launch_gather(g, a, S(stream));
}

int gcnk_gather_mask(const gcnk_graph *g, const float *in_scaled, float *out_scaled, const uint32_t *mask_bits,
                     float scale, int dim, gcnk_stream_t stream) {
    GCNK_REQUIRE(g && in_scaled && out_scaled && mask_bits && dim > 0, "bad arguments");
    GatherArgs a = {};
    a.in = in_scaled; a.out = out_scaled; a.dim = dim; a.mode = MODE_MASK; a.mask_in = mask_bits; a.scale = scale;
    return launch_gather(g, a, S(stream));
}

int gcnk_mask_row_stride_bits(int dim) { return mask_stride_bits(dim); }

int gcnk_graph_rotate(gcnk_graph *g, int col_lo, int col_hi, gcnk_stream_t stream) {
    GCNK_REQUIRE(g && !g->base && !g->seg1 && col_lo >= 0 && col_hi >= col_lo && col_hi <= g->n_cols, "needs a base graph that is not rotated yet and a column range");
    cudaStream_t st = S(stream);
    int *rot = nullptr;
    GCNK_CUDA(cudaMalloc(&rot, sizeof(int) * ((size_t)g->nnz + 4)));
    GCNK_CUDA(cudaMemsetAsync(rot + g->nnz, 0, sizeof(int) * 4, st));
    GCNK_CUDA(cudaMalloc(&g->seg1, sizeof(int) * std::max(g->n, 1)));
    GCNK_CUDA(cudaMalloc(&g->seg2, sizeof(int) * std::max(g->n, 1)));
    g->seg_owned = true;
    if (g->n) { rotate_rows_kernel<<<(g->n + 7) / 8, 256, 0, st>>>(g->indptr, g->indices, g->n, col_lo, col_hi, rot, g->seg1, g->seg2); GCNK_LAUNCHED(); }
    GCNK_CUDA(cudaStreamSynchronize(st));
    g->own_indices = rot;                 // freed with the handle; the caller's array is no longer referenced
    g->indices = rot;
    g->idx4_ok = g->nnz > 0 && idx4_readable(rot, g->nnz);
    g->rot_lo = col_lo; g->rot_hi = col_hi;
    g->symmetric = 0;                     // (a slice is never square anyway)
    return GCNK_OK;
}

int gcnk_gather_exchange_next(const float *own_rows, size_t n_floats, float *const *peer_rows, int n_peers, int *const *peer_flag_slots,
                              const int *d_wait_flags, int rank, int world, int value, unsigned *d_counter, int *d_err) {
    GCNK_REQUIRE(own_rows && peer_rows && peer_flag_slots && d_wait_flags && d_counter && d_err && n_peers >= 1 && n_peers <= 7 && world >= 2 &&
                     world <= 8 && rank >= 0 && rank < world && n_floats % 4 == 0,
                 "bad arguments");
    t_xchg.src = own_rows; t_xchg.n_floats = n_floats; t_xchg.peers = n_peers;
    for (int i = 0; i < n_peers; i++) { t_xchg.dst[i] = peer_rows[i]; t_xchg.flag[i] = peer_flag_slots[i]; }
    t_xchg.counter = d_counter; t_xchg.value = value; t_xchg.rank = rank; t_xchg.world = world; t_xchg.wait_flags = d_wait_flags; t_xchg.err = d_err;
    t_xchg.armed = true;
    return GCNK_OK;
}

int gcnk_gather_init_next(const float *d_partial) {
    t_init = d_partial;
    return GCNK_OK;
}

int gcnk_gather_raw(const gcnk_graph *g, const float *in_scaled, float *out_raw, int dim, gcnk_stream_t stream) {
    GCNK_REQUIRE(g && in_scaled && out_raw && dim > 0, "bad arguments");
    GatherArgs a = {};
    a.in = in_scaled; a.out = out_raw; a.dim = dim; a.mode = MODE_RAW; a.scale = 1.f;
    return launch_gather(g, a, S(stream));
}

int gcnk_gather_wait_next(const int *d_flags, int n_flags, int skip, int value, int *d_err) {
    GCNK_REQUIRE((d_flags && n_flags > 0 && n_flags <= THREADS && d_err) || (!d_flags && n_flags == 0), "bad arguments");
    t_wait.flags = d_flags; t_wait.n = n_flags; t_wait.skip = skip; t_wait.value = value; t_wait.err = d_err;
    return GCNK_OK;
}

int gcnk_gather_variant(int v) {
    const int before = gather_variant();
    if (v >= 0 && v <= 3) g_gather_variant = v;
    return before;
}

int gcnk_graphsum(const gcnk_graph *gc, const float *in, float *out, int dim, gcnk_stream_t stream) {
    GCNK_REQUIRE(gc && in && out && dim > 0, "bad arguments");
    gcnk_graph *g = const_cast<gcnk_graph *>(gc);
    if (g->n == 0) return GCNK_OK;
    // widths 5, 6, 7, 9, ... : through padded rows (16-byte pitch) on both sides of the gather
    const int pitch = (dim > 4 && dim % 4) ? (dim + 3) & ~3 : dim;
    const bool padded = pitch != dim;
    const size_t need_in = (size_t)g->n_cols * pitch, need = need_in + (padded ? (size_t)g->n * pitch : 0);
    if (g->scratch_elems < need) {
        GCNK_CUDA(cudaStreamSynchronize(S(stream)));
        if (g->scratch) GCNK_CUDA(cudaFree(g->scratch));
        g->scratch = nullptr; g->scratch_elems = 0;
        GCNK_CUDA(cudaMalloc(&g->scratch, sizeof(float) * std::max<size_t>(need, 4)));
        g->scratch_elems = need;
    }
    if (!padded) {
        int rc = gcnk_scale_rows(g->dinv_cols, in, g->scratch, g->n_cols, dim, stream);
        if (rc) return rc;
        return gcnk_gather_plain(g, g->scratch, out, dim, stream);
    }
    float *out_p = g->scratch + need_in;
    const int64_t total_in = (int64_t)g->n_cols * pitch, total_out = (int64_t)g->n * dim;
    const int64_t cap = (int64_t)sm_count() * 16;
    scale_rows_pad_kernel<<<(int)std::min<int64_t>((total_in + 255) / 256, cap), 256, 0, S(stream)>>>(g->dinv_cols, in, g->scratch, total_in, dim, pitch);
    GCNK_LAUNCHED();
    int rc = gcnk_gather_plain(g, g->scratch, out_p, pitch, stream);
    if (rc) return rc;
    unpad_rows_kernel<<<(int)std::min<int64_t>((total_out + 255) / 256, cap), 256, 0, S(stream)>>>(out_p, out, total_out, dim, pitch);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_graph_release_scratch(gcnk_graph *g) {
    GCNK_REQUIRE(g, "null graph");
    if (g->scratch) GCNK_CUDA(cudaFree(g->scratch));
    g->scratch = nullptr; g->scratch_elems = 0;
    return GCNK_OK;
}

int gcnk_partition_rows(const int *h_indptr, int n, int parts, int *h_row_begin) {
    GCNK_REQUIRE(h_indptr && h_row_begin && n >= 0 && parts > 0, "bad arguments");
    const int64_t nnz = h_indptr[n];
    h_row_begin[0] = 0;
    for (int k = 1; k < parts; k++) {
        // first row whose prefix nnz reaches k/parts of the total (keeps every cut monotone)
        const int64_t target = (nnz * k + parts - 1) / parts;
        const int *p = std::lower_bound(h_indptr + h_row_begin[k - 1], h_indptr + n, (int)std::min<int64_t>(target, INT32_MAX));
        // even cuts: with 64-byte rows (width 16) no 128-byte line of a gather source then holds rows of two ranks
        h_row_begin[k] = std::min(n, ((int)(p - h_indptr) + 1) & ~1);
    }
    h_row_begin[parts] = n;
    return GCNK_OK;
}

}  // extern "C"
