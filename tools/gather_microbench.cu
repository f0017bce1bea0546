// tools/gather_microbench.cu — how fast can one B200 SM fetch random 64-byte rows of an L2-resident [N x 16] fp32
// array, by which path?  ("get GraphSum off the LSU pipe — measure before rejecting".)
//
// GraphSum at hidden 16 is one random 64-byte row read per edge.  The production kernel (csrc/graph.cu) issues them as
// LDG.128 from four lanes per row and is bound by the L1TEX LSU data pipe: one wavefront per row per SM clock.  This
// program times the alternatives on the same access pattern (uniform random rows of a 14.9 MB array, summed per
// 512-edge segment like a row of the Reddit-shape graph, 64 B written per segment):
//
//   lsu      the production inner loop: aligned int4 index reads, four LDG.128 row reads in flight per lane
//   bulk     per-lane cp.async.bulk.shared.global (64 B, UBLKCP) into a warp-private shared-memory ring with
//            mbarrier complete_tx, rows then summed from CONTIGUOUS shared memory with LDS.128 (1/2 wavefront per row)
//   gather4  cp.async.bulk.tensor.2d...tile::gather4 (four row indices per instruction, 256 B, UTMALDG) into the same ring
//   lds      the consumer side alone (rows already in shared memory): the ceiling of any staged design
//   mix      a fraction of the warps runs `lsu`, the rest `bulk`, on disjoint segments — do the two paths add up?
//
// Output: one line per run with microseconds, rows per SM clock (at the clock measured by a spin kernel) and the
// gathered GB/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/build/gather_microbench
// tools/gather_microbench.cu -lcuda.   Every mbarrier wait is bounded (~1 s) so a protocol error cannot hang the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr unsigned FULL = 0xffffffffu;
constexpr int DIM = 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
__device__ __forceinline__ void bulk_row(void *dst, const void *src, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ int4 ld_idx4(const int *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void add4(float4 &a, const float4 &x) { a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w; }

// sum over the eight 4-lane groups, lanes 0..3 then hold the 16 floats of the segment
__device__ __forceinline__ void reduce_store(float4 acc, float *out, int seg, int lane) {
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
        acc.x += __shfl_xor_sync(FULL, acc.x, off); acc.y += __shfl_xor_sync(FULL, acc.y, off);
        acc.z += __shfl_xor_sync(FULL, acc.z, off); acc.w += __shfl_xor_sync(FULL, acc.w, off);
    }
    if (lane < 4) reinterpret_cast<float4 *>(out + (size_t)seg * DIM)[lane] = acc;
}

// ------------------------------------------------------------------------------------------ lsu --
// warps [w0, w0 + nw_total) of the launch (all CTAs) take segments seg0 + k*nw_total + (gw - w0)
__device__ __forceinline__ void lsu_warp(const float *__restrict__ src, const int *__restrict__ idx, float *out, int seg_first, int seg_end,
                                         int seg_step, int seg_len, int lane) {
    const int g4 = (lane >> 2) * 4, q = lane & 3;
    const float *in_q = src + q * 4;
    for (int seg = seg_first; seg < seg_end; seg += seg_step) {
        const int *ip = idx + (size_t)seg * seg_len + g4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 cur = ld_idx4(ip);
        for (int e = 0; e < seg_len; e += 32) {
            int4 nxt = cur;
            if (e + 32 < seg_len) nxt = ld_idx4(ip + e + 32);
            const float4 x0 = __ldg(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)cur.x * DIM));
            const float4 x1 = __ldg(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)cur.y * DIM));
            const float4 x2 = __ldg(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)cur.z * DIM));
            const float4 x3 = __ldg(reinterpret_cast<const float4 *>(in_q + (size_t)(unsigned)cur.w * DIM));
            add4(acc, x0); add4(acc, x1); add4(acc, x2); add4(acc, x3);
            cur = nxt;
        }
        reduce_store(acc, out, seg, lane);
    }
}

// ----------------------------------------------------------------------------------- bulk / gather4 --
// MODE 0: per-lane 64-byte bulk copies; MODE 1: gather4 from lanes 0..7; MODE 2: no fetch at all (lds ceiling)
template <int SLOTS, int MODE>
__device__ __forceinline__ void staged_warp(const float *__restrict__ src, const CUtensorMap *map, const int *__restrict__ idx, float *out,
                                            int seg_first, int seg_end, int seg_step, int seg_len, float *ring, uint64_t *bars, int lane, int *err) {
    const int cps = seg_len / 32;                               // chunks per segment
    const int nseg = seg_first < seg_end ? (seg_end - seg_first + seg_step - 1) / seg_step : 0;
    const long long total = (long long)nseg * cps;              // this warp's chunk stream
    if (lane == 0) for (int s = 0; s < SLOTS; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto chunk_ptr = [&](long long t) { return idx + (size_t)(seg_first + (int)(t / cps) * seg_step) * seg_len + (int)(t % cps) * 32; };
    auto issue = [&](long long t, int slot) {
        float *dst = ring + slot * 32 * DIM;
        if (MODE == 0) {
            const int r = __ldg(chunk_ptr(t) + lane);
            if (lane == 0) mbar_expect_tx(&bars[slot], 32 * 64);
            __syncwarp();
            bulk_row(dst + lane * DIM, src + (size_t)(unsigned)r * DIM, &bars[slot]);
        } else if (MODE == 1) {
            if (lane == 0) mbar_expect_tx(&bars[slot], 32 * 64);
            __syncwarp();
            if (lane < 8) {
                const int4 r = ld_idx4(chunk_ptr(t) + lane * 4);
                tma_gather4(dst + lane * 4 * DIM, map, &bars[slot], 0, r.x, r.y, r.z, r.w);
            }
        }
    };
    const long long pre = total < SLOTS ? total : SLOTS;
    for (long long t = 0; t < pre; t++) issue(t, (int)t);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *ring4 = reinterpret_cast<const float4 *>(ring);
    for (long long t = 0; t < total; t++) {
        const int slot = (int)(t % SLOTS);
        if (MODE != 2) { if (!mbar_wait(&bars[slot], (uint32_t)((t / SLOTS) & 1), err)) return; }
        // rows k*8 + lane/4, quarter lane%4: a warp-wide LDS.128 reads 512 contiguous bytes (4 wavefronts for 8 rows)
        const float4 x0 = ring4[slot * 128 + lane], x1 = ring4[slot * 128 + 32 + lane];
        const float4 x2 = ring4[slot * 128 + 64 + lane], x3 = ring4[slot * 128 + 96 + lane];
        add4(acc, x0); add4(acc, x1); add4(acc, x2); add4(acc, x3);
        __syncwarp();
        if (t + SLOTS < total) issue(t + SLOTS, slot);
        if ((t + 1) % cps == 0) {
            reduce_store(acc, out, seg_first + (int)(t / cps) * seg_step, lane);
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

template <int SLOTS, int MODE>
__global__ void __launch_bounds__(256) k_staged(const float *src, const __grid_constant__ CUtensorMap map, const int *idx, float *out, int nseg, int seg_len,
                                                int lsu_warps_per_cta, int seg_split, int *err) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int st_warps = nw - lsu_warps_per_cta;                // staged warps per CTA
    if (warp < lsu_warps_per_cta) {                             // mix: these warps run the LSU loop on segments [0, seg_split)
        const int gw = blockIdx.x * lsu_warps_per_cta + warp, GW = gridDim.x * lsu_warps_per_cta;
        lsu_warp(src, idx, out, gw, seg_split, GW, seg_len, lane);
        return;
    }
    const int sw = warp - lsu_warps_per_cta;
    float *ring = reinterpret_cast<float *>(smem) + (size_t)sw * SLOTS * 32 * DIM;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)st_warps * SLOTS * 32 * DIM * 4) + sw * SLOTS;
    const int gw = blockIdx.x * st_warps + sw, GW = gridDim.x * st_warps;
    staged_warp<SLOTS, MODE>(src, &map, idx, out, seg_split + gw, nseg, GW, seg_len, ring, bars, lane, err);
}

__global__ void __launch_bounds__(256) k_lsu(const float *src, const int *idx, float *out, int nseg, int seg_len) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    lsu_warp(src, idx, out, blockIdx.x * nw + warp, nseg, gridDim.x * nw, seg_len, lane);
}

__global__ void k_clock(long long *out, long long spin) {
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (threadIdx.x == 0) *out = clock64() - t0;
}

static bool make_map(CUtensorMap *map, const float *base, uint64_t rows, uint32_t box_rows) {
    typedef CUresult (*fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) return false;
    const cuuint64_t dims[2] = {DIM, rows}, strides[1] = {DIM * sizeof(float)};
    const cuuint32_t box[2] = {DIM, box_rows}, estr[2] = {1, 1};
    return reinterpret_cast<fn_t>(f)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "lsu";
    const int ctas_per_sm = argc > 2 ? atoi(argv[2]) : 4;
    const int slots = argc > 3 ? atoi(argv[3]) : 4;
    const double lsu_frac = argc > 4 ? atof(argv[4]) : 0.0;     // mix: share of the segments given to the LSU warps
    const int lsu_warps = argc > 5 ? atoi(argv[5]) : 0;         // mix: LSU warps per CTA (of 8)
    const int box_rows = argc > 6 ? atoi(argv[6]) : 1;          // gather4: rows of the tensor map's box
    const int N = 232965, seg_len = 512;
    const int nseg = 28 * 1024 * 1024 / seg_len;                // 29.4 M edges per launch (a quarter of the Reddit-shape graph)
    const size_t E = (size_t)nseg * seg_len;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    std::vector<float> h_src((size_t)N * DIM);
    std::vector<int> h_idx(E);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (auto &v : h_src) v = (float)((rnd() >> 40) & 0xffff) * (1.0f / 65536.0f);
    for (auto &v : h_idx) v = (int)(rnd() % (uint64_t)N);
    float *d_src, *d_out, *d_ref;
    int *d_idx, *d_err;
    long long *d_clk;
    CK(cudaMalloc(&d_src, h_src.size() * 4)); CK(cudaMalloc(&d_idx, E * 4 + 64)); CK(cudaMalloc(&d_out, (size_t)nseg * DIM * 4));
    CK(cudaMalloc(&d_ref, (size_t)nseg * DIM * 4)); CK(cudaMalloc(&d_err, 4)); CK(cudaMalloc(&d_clk, 8));
    CK(cudaMemcpy(d_src, h_src.data(), h_src.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_idx, h_idx.data(), E * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_err, 0, 4)); CK(cudaMemset(d_out, 0, (size_t)nseg * DIM * 4));

    CUtensorMap map;
    memset(&map, 0, sizeof map);
    const bool need_map = !strcmp(mode, "gather4");
    if (need_map && !make_map(&map, d_src, N, box_rows)) { fprintf(stderr, "cuTensorMapEncodeTiled failed (box rows %d)\n", box_rows); return 3; }

    // SM clock under load: spin for a known number of cycles, time it with events
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto sm_mhz = [&]() {
        const long long spin = 200000000LL;
        CK(cudaEventRecord(e0)); k_clock<<<sms, 32>>>(d_clk, spin); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        long long cyc; CK(cudaMemcpy(&cyc, d_clk, 8, cudaMemcpyDeviceToHost));
        return (double)cyc / (ms * 1e3);
    };

    const int grid = sms * ctas_per_sm;
    int st_mode = !strcmp(mode, "bulk") || !strcmp(mode, "mix") ? 0 : !strcmp(mode, "gather4") ? 1 : !strcmp(mode, "lds") ? 2 : -1;
    const int lw = !strcmp(mode, "mix") ? lsu_warps : 0;
    int seg_split = !strcmp(mode, "mix") ? (int)(nseg * lsu_frac) : 0;
    const size_t smem = (size_t)(8 - lw) * slots * 32 * DIM * 4 + (size_t)(8 - lw) * slots * 8;
    auto launch = [&](float *out) {
        if (st_mode < 0) { k_lsu<<<grid, 256>>>(d_src, d_idx, out, nseg, seg_len); return; }
#define LAUNCH(S, M) do { CK(cudaFuncSetAttribute(k_staged<S, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                          k_staged<S, M><<<grid, 256, smem>>>(d_src, map, d_idx, out, nseg, seg_len, lw, seg_split, d_err); } while (0)
        if (slots == 2) { if (st_mode == 0) LAUNCH(2, 0); else if (st_mode == 1) LAUNCH(2, 1); else LAUNCH(2, 2); }
        else if (slots == 3) { if (st_mode == 0) LAUNCH(3, 0); else if (st_mode == 1) LAUNCH(3, 1); else LAUNCH(3, 2); }
        else if (slots == 4) { if (st_mode == 0) LAUNCH(4, 0); else if (st_mode == 1) LAUNCH(4, 1); else LAUNCH(4, 2); }
        else if (slots == 6) { if (st_mode == 0) LAUNCH(6, 0); else if (st_mode == 1) LAUNCH(6, 1); else LAUNCH(6, 2); }
        else { if (st_mode == 0) LAUNCH(8, 0); else if (st_mode == 1) LAUNCH(8, 1); else LAUNCH(8, 2); }
    };
    k_lsu<<<sms * 6, 256>>>(d_src, d_idx, d_ref, nseg, seg_len);   // reference sums
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < 2; i++) launch(d_out);
    CK(cudaDeviceSynchronize());
    int err = 0;
    CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    if (err) { printf("{\"mode\": \"%s\", \"error\": \"mbarrier wait timed out (code %d)\"}\n", mode, err); return 4; }
    const int reps = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) launch(d_out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps, mhz = sm_mhz();
    // check against the LSU sums (same additions in a different order: compare with a tolerance)
    std::vector<float> a((size_t)nseg * DIM), b((size_t)nseg * DIM);
    CK(cudaMemcpy(a.data(), d_out, a.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), d_ref, b.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    if (st_mode != 2) for (size_t i = 0; i < a.size(); i++) { const double d = fabs((double)a[i] - b[i]) / (fabs((double)b[i]) + 1e-6); if (d > worst) worst = d; }
    printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"slots\": %d, \"lsu_frac\": %.2f, \"lsu_warps\": %d, \"box_rows\": %d, \"edges\": %zu, \"us\": %.1f, "
           "\"sm_mhz\": %.0f, \"rows_per_clk_per_sm\": %.3f, \"gathered_GBps\": %.0f, \"max_rel_diff_vs_lsu\": %.2e}\n",
           mode, ctas_per_sm, slots, lsu_frac, lw, box_rows, E, us, mhz, (double)E / sms / (us * mhz), (double)E * 64 / us * 1e-3, worst);
    return 0;
}
