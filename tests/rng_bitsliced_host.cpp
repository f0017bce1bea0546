// Host check of cuda_gcn_b200/csrc/rng_bitsliced.cuh: the per-thread bit-sliced generator against the scalar xorshift128+
// stream (reference: src/seq/rand.cpp:17-28, keep rule module.cpp:211-216).  Built and run by tests/test_host_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../cuda_gcn_b200/csrc/rng_bitsliced.cuh"

using namespace gcnk_bs;

static State128 step(State128 s) {
    uint64_t t = s.lo;
    const uint64_t u = s.hi;
    t ^= t << 23;
    t ^= t >> 17;
    t ^= u ^ (u >> 26);
    return State128{u, t};
}
static State128 apply(const State128 *cols, State128 s) {
    State128 r{0, 0};
    for (int j = 0; j < 64; j++) if ((s.lo >> j) & 1) { r.lo ^= cols[j].lo; r.hi ^= cols[j].hi; }
    for (int j = 0; j < 64; j++) if ((s.hi >> j) & 1) { r.lo ^= cols[64 + j].lo; r.hi ^= cols[64 + j].hi; }
    return r;
}

template <int LS>
static int check(State128 base, int threshold, int64_t left, const Entry *tab) {
    constexpr int64_t TOTAL = 32ll << LS;
    std::vector<uint32_t> got((size_t)(TOTAL / 32) + 1, 0xdeadbeefu), want((size_t)(TOTAL / 32) + 1, 0xdeadbeefu);
    if (threshold == 0x40000000) generate<LS, true>(base, tab, threshold, got.data(), left);
    else generate<LS, false>(base, tab, threshold, got.data(), left);
    State128 s = base;
    const int64_t n = left < TOTAL ? left : TOTAL;
    for (int64_t i = 0; i < n; i++) {
        s = step(s);
        const int draw = (int)((s.hi + s.lo) & 0x7fffffffull);
        if (i % 32 == 0) want[(size_t)(i / 32)] = 0;
        want[(size_t)(i / 32)] |= (uint32_t)(draw >= threshold) << (i % 32);
    }
    for (size_t w = 0; w < got.size(); w++)
        if (got[w] != want[w]) {
            printf("LS=%d thr=%d left=%lld: word %zu got %08x want %08x\n", LS, threshold, (long long)left, w, got[w], want[w]);
            return 1;
        }
    return 0;
}

int main() {
    // transpose: bit k of A'[i] == bit i of A[k]
    uint32_t A[32], B[32];
    srand(7);
    for (int k = 0; k < 32; k++) A[k] = B[k] = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    transpose32(B);
    for (int i = 0; i < 32; i++)
        for (int k = 0; k < 32; k++)
            if (((B[i] >> k) & 1u) != ((A[k] >> i) & 1u)) { printf("transpose wrong at %d,%d\n", i, k); return 1; }

    // M^(2^b) columns by squaring
    static State128 J[11][128];
    for (int j = 0; j < 128; j++) J[0][j] = step(State128{j < 64 ? 1ull << j : 0, j >= 64 ? 1ull << (j - 64) : 0});
    for (int b = 1; b <= 10; b++) for (int j = 0; j < 128; j++) J[b][j] = apply(J[b - 1], J[b - 1][j]);
    static Entry tab7[NIB_ENTRIES], tab10[NIB_ENTRIES], tab5[NIB_ENTRIES];
    build_nibble_tables(J[7], tab7);
    build_nibble_tables(J[10], tab10);
    build_nibble_tables(J[5], tab5);

    // nibble application == column application
    State128 s{0x123456789abcdef0ull, 0x0fedcba987654321ull};
    uint32_t w[4] = {(uint32_t)s.lo, (uint32_t)(s.lo >> 32), (uint32_t)s.hi, (uint32_t)(s.hi >> 32)};
    apply_nibbles(tab7, w);
    const State128 r = apply(J[7], s);
    if (w[0] != (uint32_t)r.lo || w[1] != (uint32_t)(r.lo >> 32) || w[2] != (uint32_t)r.hi || w[3] != (uint32_t)(r.hi >> 32)) { printf("nibble apply wrong\n"); return 1; }

    int bad = 0;
    const int thresholds[] = {0, 1, 0x40000000, (int)(0.9f * (float)0x7fffffff), (int)(0.1f * (float)0x7fffffff), 0x7fffffff, 123456789};
    const State128 bases[] = {{1804289383ull, 846930886ull}, {0xffffffffffffffffull, 1ull}, {0x9e3779b97f4a7c15ull, 0xbf58476d1ce4e5b9ull}};
    for (const State128 &b0 : bases)
        for (int thr : thresholds) {
            for (int64_t left : {1ll, 31ll, 32ll, 33ll, 127ll, 128ll, 129ll, 1000ll, 1023ll, 1024ll, 1025ll, 4095ll, 4096ll, 5000ll}) {
                bad += check<7>(b0, thr, left, tab7);
                bad += check<5>(b0, thr, left, tab5);
            }
            for (int64_t left : {1ll, 1023ll, 1024ll, 1025ll, 20000ll, 32767ll, 32768ll, 40000ll}) bad += check<10>(b0, thr, left, tab10);
        }
    if (bad) { printf("%d mismatches\n", bad); return 1; }
    printf("ok\n");
    return 0;
}
