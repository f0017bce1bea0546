// rng.cu — the reference's xorshift128+ stream (src/seq/rand.cpp:17-28), reproduced bit-for-bit in
// parallel, and the Dropout keep-bit generator built on it (semantics src/seq/module.cpp:207-221).
// Replaces cuda_Dropout_forward_kernel's 1,024 shared cuRAND states (src/cuda/cuda_kernel.cu:223-234,
// a data race) and cuda_init_rand_kernel (:244-248).
//
// xorshift128+ advances its 128-bit state by a map that is linear over GF(2):
//     (s0, s1) -> (s1, t ^ (t<<23) ^ ((t ^ (t<<23)) >> 17) ^ s1 ^ (s1>>26)),  t = s0.
// So "advance by k draws" is a 128x128 bit matrix M^k.  The powers J_b = M^(2^b), b = 0..63, are built
// once on the host (63 squarings) and kept on the device (128 KB).  A CTA of 128 threads owns 2^16
// consecutive draws: it jumps from the stream position to its own offset cooperatively (thread i owns
// state bit i, so one J_b application is a 128-way XOR reduction), then fans out to the 128 per-thread
// start states by doubling (state[t + 2^b] = J_(9+b) * state[t]), and every thread runs the plain
// sequential generator for its 512 draws, emitting one keep bit per draw.  The draw sequence — and with
// it every dropout mask — is therefore identical to gcn-seq's for the same seed, at about the cost of
// a counter-based generator, with no per-thread state kept between launches.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "rng_bitsliced.cuh"

using namespace gcnk;

namespace {

struct U128 { uint64_t lo, hi; };   // lo = s0 (bits 0..63), hi = s1 (bits 64..127)

__host__ __device__ inline U128 step_state(U128 s) {
    uint64_t t = s.lo;
    const uint64_t u = s.hi;
    t ^= t << 23;
    t ^= t >> 17;
    t ^= u ^ (u >> 26);
    return U128{u, t};
}

constexpr int N_POW = 64;
constexpr int CTA_THREADS = 128;   // a CTA owns 128 << TS consecutive draws, TS = log2(draws per thread): 9 for long streams, 7 for short ones

struct JumpTables {
    // J[b][j] = M^(2^b) applied to the unit vector e_j  (column j of the matrix)
    U128 J[N_POW][128];
};

U128 apply_host(const U128 *cols, U128 s) {
    U128 r{0, 0};
    for (int j = 0; j < 64; j++) if ((s.lo >> j) & 1) { r.lo ^= cols[j].lo; r.hi ^= cols[j].hi; }
    for (int j = 0; j < 64; j++) if ((s.hi >> j) & 1) { r.lo ^= cols[64 + j].lo; r.hi ^= cols[64 + j].hi; }
    return r;
}

const JumpTables &host_tables() {
    static JumpTables *T = nullptr;
    if (!T) {
        T = new JumpTables;
        for (int j = 0; j < 128; j++) {
            U128 e{j < 64 ? (1ull << j) : 0, j >= 64 ? (1ull << (j - 64)) : 0};
            T->J[0][j] = step_state(e);
        }
        for (int b = 1; b < N_POW; b++)
            for (int j = 0; j < 128; j++) T->J[b][j] = apply_host(T->J[b - 1], T->J[b - 1][j]);
    }
    return *T;
}

U128 skip_host(U128 s, uint64_t n) {
    const JumpTables &T = host_tables();
    for (int b = 0; b < N_POW; b++) if ((n >> b) & 1) s = apply_host(T.J[b], s);
    return s;
}

// thread-serial application (used for the in-CTA fan-out)
__device__ __forceinline__ U128 apply_dev(const U128 *__restrict__ cols, U128 s) {
    uint64_t lo = 0, hi = 0;
#pragma unroll 8
    for (int j = 0; j < 64; j++) {
        const uint64_t mk = 0 - ((s.lo >> j) & 1ull);
        const ulonglong2 c = *reinterpret_cast<const ulonglong2 *>(cols + j);
        lo ^= c.x & mk; hi ^= c.y & mk;
    }
#pragma unroll 8
    for (int j = 0; j < 64; j++) {
        const uint64_t mk = 0 - ((s.hi >> j) & 1ull);
        const ulonglong2 c = *reinterpret_cast<const ulonglong2 *>(cols + 64 + j);
        lo ^= c.x & mk; hi ^= c.y & mk;
    }
    return U128{lo, hi};
}

template <int THREAD_SHIFT>
__global__ void __launch_bounds__(CTA_THREADS) dropout_mask_kernel(const U128 *__restrict__ J, U128 start, uint32_t *__restrict__ keep,
                                                                    int64_t n, int threshold) {
    constexpr int DRAWS_PER_THREAD = 1 << THREAD_SHIFT, CTA_SHIFT = THREAD_SHIFT + 7;
    __shared__ U128 s_state[CTA_THREADS];
    __shared__ uint64_t s_red[2][CTA_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // 1) CTA start state = M^(blockIdx << 16) * start, cooperatively: thread `tid` owns state bit `tid`
    U128 st = start;
    for (unsigned b = 0, blk = blockIdx.x; blk; b++, blk >>= 1) {
        if (!(blk & 1)) continue;
        const bool bit = tid < 64 ? (st.lo >> tid) & 1 : (st.hi >> (tid - 64)) & 1;
        const U128 col = J[(CTA_SHIFT + b) * 128 + tid];
        uint64_t lo = bit ? col.lo : 0, hi = bit ? col.hi : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo ^= __shfl_xor_sync(FULL, lo, o); hi ^= __shfl_xor_sync(FULL, hi, o); }
        if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
        __syncthreads();
        st.lo = s_red[0][0] ^ s_red[0][1] ^ s_red[0][2] ^ s_red[0][3];
        st.hi = s_red[1][0] ^ s_red[1][1] ^ s_red[1][2] ^ s_red[1][3];
        __syncthreads();
    }

    // 2) fan out to per-thread start states by doubling: state[t + 2^b] = J_(THREAD_SHIFT+b) * state[t]
    if (tid == 0) s_state[0] = st;
    __syncthreads();
    for (int b = 0; (1 << b) < CTA_THREADS; b++) {
        const int have = 1 << b;
        // only threads whose draws exist need a state
        if (tid < have && (((int64_t)blockIdx.x << CTA_SHIFT) + ((int64_t)(tid + have) << THREAD_SHIFT)) < n)
            s_state[tid + have] = apply_dev(J + (THREAD_SHIFT + b) * 128, s_state[tid]);
        __syncthreads();
    }
    const int64_t first = ((int64_t)blockIdx.x << CTA_SHIFT) + ((int64_t)tid << THREAD_SHIFT);
    if (first >= n) return;
    U128 s = s_state[tid];

    // 3) the sequential generator: 512 draws -> 16 keep words, stored as four 16-byte vectors
    uint32_t *out = keep + (first >> 5);
    const int64_t left = n - first;
#pragma unroll 1
    for (int w4 = 0; w4 < DRAWS_PER_THREAD / 128; w4++) {
        uint32_t word[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t bits = 0;
#pragma unroll
            for (int i = 0; i < 32; i++) {
                s = step_state(s);
                const uint32_t draw = (uint32_t)((s.hi + s.lo) & 0x7fffffffull);   // (t + s) & 0x7fffffff (rand.cpp:26)
                bits |= (uint32_t)((int)draw >= threshold) << i;
            }
            word[w] = bits;
        }
        const int64_t done = (int64_t)w4 * 128;
        if (left >= done + 128) {
            *reinterpret_cast<uint4 *>(out + w4 * 4) = make_uint4(word[0], word[1], word[2], word[3]);
        } else {
            for (int w = 0; w < 4; w++) {
                const int64_t lo = done + w * 32;
                if (lo >= left) break;
                const int64_t valid = left - lo;
                out[w4 * 4 + w] = valid >= 32 ? word[w] : (word[w] & ((1u << valid) - 1u));   // bits past n stay 0
            }
            return;
        }
    }
}

U128 *device_tables() {
    static U128 *d_tab[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (!d_tab[dev]) {
        const JumpTables &T = host_tables();
        if (cudaMalloc(&d_tab[dev], sizeof(JumpTables)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d_tab[dev], &T, sizeof(JumpTables), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    }
    return d_tab[dev];
}

// The bit-sliced form (rng_bitsliced.cuh): a thread runs 32 streams of 2^LS draws (one keep word per step for all of them),
// ~9 instructions per draw instead of ~24.  Start states: the same cooperative jump to the CTA's offset and doubling fan-out
// to the threads as above (a thread owns 2^(LS+5) draws), then the in-thread chain of jumps by 2^LS.
template <int LS, bool HALF>
__global__ void __launch_bounds__(CTA_THREADS) dropout_mask_bs_kernel(const U128 *__restrict__ J, const gcnk_bs::Entry *__restrict__ nib, U128 start,
                                                                       uint32_t *__restrict__ keep, int64_t n, int threshold) {
    constexpr int THREAD_SHIFT = LS + 5, CTA_SHIFT = THREAD_SHIFT + 7;
    __shared__ U128 s_state[CTA_THREADS];
    __shared__ uint64_t s_red[2][CTA_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    U128 st = start;
    for (unsigned b = 0, blk = blockIdx.x; blk; b++, blk >>= 1) {
        if (!(blk & 1)) continue;
        const bool bit = tid < 64 ? (st.lo >> tid) & 1 : (st.hi >> (tid - 64)) & 1;
        const U128 col = J[(CTA_SHIFT + b) * 128 + tid];
        uint64_t lo = bit ? col.lo : 0, hi = bit ? col.hi : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo ^= __shfl_xor_sync(FULL, lo, o); hi ^= __shfl_xor_sync(FULL, hi, o); }
        if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
        __syncthreads();
        st.lo = s_red[0][0] ^ s_red[0][1] ^ s_red[0][2] ^ s_red[0][3];
        st.hi = s_red[1][0] ^ s_red[1][1] ^ s_red[1][2] ^ s_red[1][3];
        __syncthreads();
    }
    if (tid == 0) s_state[0] = st;
    __syncthreads();
    for (int b = 0; (1 << b) < CTA_THREADS; b++) {
        const int have = 1 << b;
        if (tid < have && (((int64_t)blockIdx.x << CTA_SHIFT) + ((int64_t)(tid + have) << THREAD_SHIFT)) < n)
            s_state[tid + have] = apply_dev(J + (THREAD_SHIFT + b) * 128, s_state[tid]);
        __syncthreads();
    }
    const int64_t first = ((int64_t)blockIdx.x << CTA_SHIFT) + ((int64_t)tid << THREAD_SHIFT);
    if (first >= n) return;
    const U128 s = s_state[tid];
    gcnk_bs::generate<LS, HALF>(gcnk_bs::State128{s.lo, s.hi}, nib, threshold, keep + (first >> 5), n - first);
}

// nibble tables of M^(2^LS), LS = 7..10 (the in-thread jump chain of the bit-sliced kernels), 8 KB each
constexpr int BS_LS_MIN = 7, BS_LS_MAX = 10;
gcnk_bs::Entry *device_nibble_tables() {
    static gcnk_bs::Entry *d_tab[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (!d_tab[dev]) {
        const JumpTables &T = host_tables();
        static_assert(sizeof(gcnk_bs::State128) == sizeof(U128), "same layout");
        std::vector<gcnk_bs::Entry> h((BS_LS_MAX - BS_LS_MIN + 1) * gcnk_bs::NIB_ENTRIES);
        for (int ls = BS_LS_MIN; ls <= BS_LS_MAX; ls++)
            gcnk_bs::build_nibble_tables(reinterpret_cast<const gcnk_bs::State128 *>(T.J[ls]), h.data() + (ls - BS_LS_MIN) * gcnk_bs::NIB_ENTRIES);
        if (cudaMalloc(&d_tab[dev], sizeof(gcnk_bs::Entry) * h.size()) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d_tab[dev], h.data(), sizeof(gcnk_bs::Entry) * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    }
    return d_tab[dev];
}

template <int LS>
void launch_bs(const U128 *tab, const gcnk_bs::Entry *nib, U128 start, uint32_t *keep, int64_t n, int threshold, cudaStream_t st) {
    const unsigned ctas = (unsigned)((n + (1ll << (LS + 12)) - 1) >> (LS + 12));
    const gcnk_bs::Entry *t = nib + (LS - BS_LS_MIN) * gcnk_bs::NIB_ENTRIES;
    if (threshold == 0x40000000) dropout_mask_bs_kernel<LS, true><<<ctas, CTA_THREADS, 0, st>>>(tab, t, start, keep, n, threshold);   // dropout 0.5: the keep bit is bit 30 of the draw
    else dropout_mask_bs_kernel<LS, false><<<ctas, CTA_THREADS, 0, st>>>(tab, t, start, keep, n, threshold);
}

int rng_variant() {      // GCN_RNG_SCALAR=1: the scalar kernels only
    const char *e = getenv("GCN_RNG_SCALAR");
    return e && *e && strcmp(e, "0") ? 0 : 1;
}

}  // namespace

struct gcnk_rng { U128 s; };

extern "C" {

int gcnk_rng_create(gcnk_rng **rng, uint64_t s0, uint64_t s1) {
    GCNK_REQUIRE(rng, "null");
    *rng = new gcnk_rng{U128{s0, s1}};
    return GCNK_OK;
}
int gcnk_rng_destroy(gcnk_rng *rng) { delete rng; return GCNK_OK; }

int gcnk_rng_seed(gcnk_rng *rng, long seed) {
    GCNK_REQUIRE(rng, "null");
    srand((unsigned)seed);                       // glibc rand(), exactly as init_rand_state (rand.cpp:6-15)
    int x = 0, y = 0;
    while (x == 0 || y == 0) { x = rand(); y = rand(); }
    rng->s = U128{(uint64_t)x, (uint64_t)y};
    return GCNK_OK;
}

int gcnk_rng_get_state(const gcnk_rng *rng, uint64_t *out) { GCNK_REQUIRE(rng && out, "null"); out[0] = rng->s.lo; out[1] = rng->s.hi; return GCNK_OK; }
int gcnk_rng_set_state(gcnk_rng *rng, uint64_t s0, uint64_t s1) { GCNK_REQUIRE(rng, "null"); rng->s = U128{s0, s1}; return GCNK_OK; }
int gcnk_rng_skip(gcnk_rng *rng, uint64_t n) { GCNK_REQUIRE(rng, "null"); rng->s = skip_host(rng->s, n); return GCNK_OK; }

int gcnk_rng_next_host(gcnk_rng *rng, uint32_t *out, int64_t n) {
    GCNK_REQUIRE(rng && (out || n == 0) && n >= 0, "bad arguments");
    U128 s = rng->s;
    for (int64_t i = 0; i < n; i++) {
        s = step_state(s);
        out[i] = (uint32_t)((s.hi + s.lo) & 0x7fffffffull);
    }
    rng->s = s;
    return GCNK_OK;
}

int gcnk_dropout_mask(gcnk_rng *rng, uint32_t *keep_bits, int64_t n, float p, gcnk_stream_t stream) {
    GCNK_REQUIRE(rng && keep_bits && n >= 0, "bad arguments");
    if (n == 0) return GCNK_OK;
    U128 *tab = device_tables();
    if (!tab) return cuda_fail(cudaGetLastError(), "xorshift jump tables", __FILE__, __LINE__);
    const int threshold = (int)(p * (float)0x7fffffff);                    // int(p * MY_RAND_MAX) (module.cpp:211)
    // short streams (the N x hidden mask: 3.7 M draws at Reddit shape) get 128 draws per thread so that they still
    // fill the machine; long ones 512 (the per-CTA jump is amortised over more draws)
    if (rng_variant() == 1 && n >= (1ll << 20)) {
        // from 2^20 draws on: the bit-sliced generator (rng_bitsliced.cuh)
        gcnk_bs::Entry *nib = device_nibble_tables();
        if (!nib) return cuda_fail(cudaGetLastError(), "xorshift nibble tables", __FILE__, __LINE__);
        // draws per stream: long streams amortise the jump chain (6.5 instructions per draw at 2^10, 8.4 at 2^7) but leave
        // fewer, longer-running CTAs (a CTA owns 2^(LS+12) draws); GCN_RNG_LS overrides
        // measured in the Reddit-shape step (r02w): 2^7 403, 2^8 405, 2^9 408, 2^10 396 epochs/s
        int ls = n >= (1ll << 25) ? 9 : 7;
        if (const char *e = getenv("GCN_RNG_LS")) { const int v = atoi(e); if (v >= BS_LS_MIN && v <= BS_LS_MAX) ls = v; }
        GCNK_REQUIRE(((n + (1ll << (ls + 12)) - 1) >> (ls + 12)) <= 0x7fffffff, "too many draws for one launch");
        if (ls == 7) launch_bs<7>(tab, nib, rng->s, keep_bits, n, threshold, S(stream));
        else if (ls == 8) launch_bs<8>(tab, nib, rng->s, keep_bits, n, threshold, S(stream));
        else if (ls == 9) launch_bs<9>(tab, nib, rng->s, keep_bits, n, threshold, S(stream));
        else launch_bs<10>(tab, nib, rng->s, keep_bits, n, threshold, S(stream));
    } else if (n < (64ll << 20)) {
        const int64_t ctas = (n + (1ll << 14) - 1) >> 14;
        prefer_carveout(dropout_mask_kernel<7>);
        dropout_mask_kernel<7><<<(unsigned)ctas, CTA_THREADS, 0, S(stream)>>>(tab, rng->s, keep_bits, n, threshold);
    } else {
        const int64_t ctas = (n + (1ll << 16) - 1) >> 16;
        GCNK_REQUIRE(ctas <= 0x7fffffff, "too many draws for one launch");
        prefer_carveout(dropout_mask_kernel<9>);
        dropout_mask_kernel<9><<<(unsigned)ctas, CTA_THREADS, 0, S(stream)>>>(tab, rng->s, keep_bits, n, threshold);
    }
    GCNK_LAUNCHED();
    rng->s = skip_host(rng->s, (uint64_t)n);
    return GCNK_OK;
}

}  // extern "C"
