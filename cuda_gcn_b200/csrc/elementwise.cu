// elementwise.cu — ReLU, Dropout application, masked labels, softmax cross-entropy (+accuracy), Adam,
// L2 penalty.  Replaces reference kernels K8-K13, K15, K16 and the thrust reductions T1/T2
// (src/cuda/cuda_kernel.cu:166-240,270-288; cuda_module.cu:121-146; cuda_gcn.cu:100-134).
// CPU semantics: src/seq/module.cpp:124-233, src/seq/optim.cpp:24-37, src/seq/gcn.cpp:78-105.
//
// All reductions are two-level with a fixed order (per-CTA tree, then CTA order), so loss, counts and
// the L2 penalty are reproducible run to run.  Adam keeps the reference's mixed fp32/fp64 arithmetic
// with explicitly rounded operations (no FMA contraction), so given bit-identical gradients its
// update is bit-identical to the CPU reference.
#include <algorithm>

#include "common.cuh"

using namespace gcnk;

namespace {

// ------------------------------------------------------------------------------------ ReLU ----
__global__ void relu_fw_kernel(float *__restrict__ x, uint32_t *__restrict__ mask, int64_t n, int training) {
    // one mask word (32 elements) per thread: no two threads ever share a word
    const int64_t words = (n + 31) / 32;
    int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; w < words; w += stride) {
        uint32_t bits = 0;
        const int64_t base = w * 32;
        const int cnt = (int)min((int64_t)32, n - base);
        for (int i = 0; i < cnt; i++) {
            const float v = x[base + i];
            const bool keep = v > 0.f;                        // NaN and -0 go to 0 (module.cpp:179-181)
            bits |= (uint32_t)keep << i;
            if (!keep) x[base + i] = 0.f;
        }
        if (training && mask) mask[w] = bits;
    }
}

__global__ void relu_bw_kernel(float *__restrict__ grad, const uint32_t *__restrict__ mask, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        if (!((mask[i >> 5] >> (i & 31)) & 1u)) grad[i] = 0.f;
}

// --------------------------------------------------------------------------------- Dropout ----
__global__ void dropout_apply_kernel(float *__restrict__ x, const uint32_t *__restrict__ keep, int64_t n, float scale) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] *= ((keep[i >> 5] >> (i & 31)) & 1u) ? scale : 0.f;
}

__global__ void set_truth_kernel(int *__restrict__ truth, const int *__restrict__ split, const int *__restrict__ label,
                                 int cur, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) truth[i] = split[i] == cur ? label[i] : -1;
}

// ------------------------------------------------------------------- softmax cross-entropy ----
struct CePartial { float loss; int count; int wrong; int pad; };

__device__ __forceinline__ void block_reduce_ce(float loss, int count, int wrong, CePartial *out) {
    __shared__ float s_loss[32];
    __shared__ int s_count[32], s_wrong[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    loss = warp_sum(loss); count = warp_sum_int(count); wrong = warp_sum_int(wrong);
    if (lane == 0) { s_loss[warp] = loss; s_count[warp] = count; s_wrong[warp] = wrong; }
    __syncthreads();
    if (warp == 0) {
        loss = lane < nwarps ? s_loss[lane] : 0.f;
        count = lane < nwarps ? s_count[lane] : 0;
        wrong = lane < nwarps ? s_wrong[lane] : 0;
        loss = warp_sum(loss); count = warp_sum_int(count); wrong = warp_sum_int(wrong);
        if (lane == 0) { out->loss = loss; out->count = count; out->wrong = wrong; out->pad = 0; }
    }
}

// thread per row: max, in-place shift, sum of exp, loss term, un-normalised gradient, accuracy flag
__global__ void __launch_bounds__(256) ce_rows_kernel(float *__restrict__ logits, const int *__restrict__ truth,
                                                       float *__restrict__ grad, int n, int c, int training,
                                                       CePartial *__restrict__ partials) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f;
    int count = 0, wrong = 0;
    if (i < n) {
        float *row = logits + (size_t)i * c;
        float *g = grad ? grad + (size_t)i * c : nullptr;
        const int t = truth[i];
        if (t < 0) {
            if (training && g) for (int j = 0; j < c; j++) g[j] = 0.f;   // zero_grad (module.cpp:129)
        } else {
            count = 1;
            float mx = -1e30f;
            for (int j = 0; j < c; j++) mx = fmaxf(mx, row[j]);
            float sum = 0.f;
            for (int j = 0; j < c; j++) {
                const float v = row[j] - mx;
                row[j] = v;                                               // in place (module.cpp:140)
                sum += expf(v);
            }
            const float tv = row[t];
            loss = logf(sum) - tv;
            bool w = false;
            for (int j = 0; j < c; j++) w |= row[j] > tv;                  // strict: ties are correct (gcn.cpp:88-93)
            wrong = w;
            if (training && g) {
                for (int j = 0; j < c; j++) g[j] = expf(row[j]) / sum;
                g[t] -= 1.0f;
            }
        }
    }
    block_reduce_ce(loss, count, wrong, partials + blockIdx.x);
}

// single CTA: sums the per-CTA partials in CTA order per thread, then a fixed tree
__global__ void __launch_bounds__(256) ce_finish_kernel(const CePartial *__restrict__ partials, int parts,
                                                         gcnk_ce_result *__restrict__ result) {
    float loss = 0.f;
    int count = 0, wrong = 0;
    for (int b = threadIdx.x; b < parts; b += blockDim.x) { loss += partials[b].loss; count += partials[b].count; wrong += partials[b].wrong; }
    __shared__ CePartial total;
    block_reduce_ce(loss, count, wrong, &total);
    __syncthreads();
    if (threadIdx.x == 0) {
        result->loss = total.loss / (float)total.count;                   // count == 0 -> NaN, as the reference
        result->count = total.count; result->wrong = total.wrong; result->pad = 0;
    }
}

__global__ void ce_scale_grad_kernel(float *__restrict__ grad, int64_t n, const gcnk_ce_result *__restrict__ result) {
    const float cnt = (float)result->count;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) grad[i] = grad[i] / cnt;                   // a true division, as module.cpp:156-158
}

__global__ void __launch_bounds__(256) accuracy_kernel(const float *__restrict__ logits, const int *__restrict__ truth,
                                                        int n, int c, CePartial *__restrict__ partials) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int count = 0, wrong = 0;
    if (i < n && truth[i] >= 0) {
        const float *row = logits + (size_t)i * c;
        const float tv = row[truth[i]];
        bool w = false;
        for (int j = 0; j < c; j++) w |= row[j] > tv;
        count = 1; wrong = w;
    }
    block_reduce_ce(0.f, count, wrong, partials + blockIdx.x);
}

__global__ void __launch_bounds__(256) accuracy_finish_kernel(const CePartial *__restrict__ partials, int parts, int *out2) {
    int count = 0, wrong = 0;
    for (int b = threadIdx.x; b < parts; b += blockDim.x) { count += partials[b].count; wrong += partials[b].wrong; }
    __shared__ CePartial total;
    block_reduce_ce(0.f, count, wrong, &total);
    __syncthreads();
    if (threadIdx.x == 0) { out2[0] = total.wrong; out2[1] = total.count; }
}

// ------------------------------------------------------------------------------------ Adam ----
constexpr int ADAM_MAX = 8;
struct AdamPack { gcnk_adam_tensor t[ADAM_MAX]; };

__global__ void __launch_bounds__(256) adam_kernel(const AdamPack pack, float step_size, float beta1, float beta2,
                                                    float eps, float weight_decay) {
    const gcnk_adam_tensor T = pack.t[blockIdx.y];
    const double ob1 = 1.0 - (double)beta1, ob2 = 1.0 - (double)beta2;     // (1.0 - beta) promotes to double (optim.cpp:32-33)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T.size; i += gridDim.x * blockDim.x) {
        float g = T.grad[i];
        const float w = T.data[i];
        if (T.decay) g = __fadd_rn(g, __fmul_rn(weight_decay, w));
        const float m = (float)__dadd_rn((double)__fmul_rn(beta1, T.m[i]), __dmul_rn(ob1, (double)g));
        const float v = (float)__dadd_rn((double)__fmul_rn(beta2, T.v[i]), __dmul_rn(__dmul_rn(ob2, (double)g), (double)g));
        T.m[i] = m;
        T.v[i] = v;
        T.data[i] = __fsub_rn(w, __fdiv_rn(__fmul_rn(step_size, m), __fadd_rn(__fsqrt_rn(v), eps)));
    }
}

// single CTA, fixed order: per-thread strided partial, then a tree
__global__ void __launch_bounds__(1024) sum_squares_kernel(const float *__restrict__ w, int64_t n, float *__restrict__ out) {
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(w[i], w[i], s);
    __shared__ float sh[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) *out = s;
    }
}

int ew_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16)); }


// ---------------------------------------------------------------- sequential-order fp32 summation ----
// The reference adds the per-row loss terms into ONE float in row order (module.cpp:125-143: `total_loss += ...`).
// With 153,756 labelled rows whose terms are all close to ln(41) in the first epochs, that sum carries a systematic
// rounding error of ~1e-4 relative (every addition rounds the term to a multiple of ulp(total) the same way) —
// more than the 1e-4 parity tolerance, so a more accurate tree sum does NOT match gcn-seq at the headline size.
// This kernel reproduces the reference's result bit for bit, without the 153,756-deep dependent chain:
//   while the running sum S stays inside one binade [2^e, 2^(e+1)), S is a multiple of u = ulp(S) = 2^(e-23) and
//   fl(S + t) = S + u*rint(t/u) for every t that is not an exact tie — independent of S.  So a block of terms is
//   rounded in parallel (t/u is an exact power-of-two scaling), summed exactly as integers (REDUX), and added to the
//   mantissa of S in integer arithmetic.  A block that would leave the binade, contains a tie, a negative or huge
//   term, or starts from S = 0 is replayed with plain sequential additions, 32 terms at a time.
// seq_fast_block is that step for one warp (the fallback granularity: 1,024 terms, then rows of 32); seq_sum_kernel spreads
// the rounding over eight warps.
__device__ __forceinline__ bool seq_fast_block(float &S, const float *t, int cnt) {
    // cnt terms per lane (lane-major inside a row is irrelevant here: the fast path is order-independent)
    const uint32_t sb = __float_as_uint(S);
    const int e = (int)((sb >> 23) & 0xff);                     // biased exponent
    bool ok = !(sb >> 31) && e >= 30 && e <= 250;               // positive, normal, and u = 2^(e-150) is a normal float too
    const float inv_u = __uint_as_float((uint32_t)(127 + 150 - e) << 23);     // 2^(150 - e) = 1 / u  (e in [30,250] -> exponent in [27,247])
    int r_sum = 0;
    for (int j = 0; j < cnt; j++) {
        const float x = t[j] * inv_u;                           // exact (power-of-two scaling) unless it over/underflows
        const float r = rintf(x);
        // ties, negative contributions, terms of the size of S itself, NaN/Inf: let the sequential path decide
        if (!(x >= 0.f) || !(x < 1048576.f) || fabsf(x - r) == 0.5f) ok = false;
        r_sum += (int)r;
    }
    ok = __all_sync(FULL, ok);
    if (!ok) return false;
    const int R = __reduce_add_sync(FULL, r_sum);               // < 2^20 * 1,024: exact
    const uint32_t m = (sb & 0x7fffffu) | 0x800000u;            // S = m * u
    if (m + (uint32_t)R >= 0x1000000u) return false;            // would cross into the next binade
    S = __uint_as_float((sb & 0x7f800000u) | ((m + (uint32_t)R) & 0x7fffffu));
    return true;
}

// Per-lane part of a fast block with the exponent given: the integer sum of the lane's rounded terms; ok = false if any of
// them is a tie, negative, as large as S itself or not a number (then the sequential path decides).
__device__ __forceinline__ int seq_round_terms(int e, const float *t, int cnt, bool &ok) {
    const float inv_u = __uint_as_float((uint32_t)(127 + 150 - e) << 23);     // 2^(150 - e) = 1 / ulp(S)
    int r_sum = 0;
    for (int j = 0; j < cnt; j++) {
        const float x = t[j] * inv_u;
        const float r = rintf(x);
        if (!(x >= 0.f) || !(x < 1048576.f) || fabsf(x - r) == 0.5f) ok = false;
        r_sum += (int)r;
    }
    return r_sum;
}

// One CTA of eight warps.  Per chunk of 4,096 terms (shared-memory ring, filled with 16-byte loads one chunk ahead):
//   all warps   round the chunk's four 1,024-term blocks in parallel, SPECULATING that the running sum stays in the binade it
//               is in at the start of the chunk (the rounding unit only depends on that exponent) — two warps per block;
//   warp 0      then walks the four blocks in order: a block whose speculation holds (same exponent, every term fine, no carry
//               out of the mantissa) is one integer addition; any other block is replayed from shared memory with the
//               row-wise fast path / the reference's own scalar loop.
// The serial part is ~40 instructions per 1,024 terms; a lone warp doing the rounding as well was issue-bound at 131 us for
// 153,756 terms (r02x).
constexpr int SEQ_CHUNK = 4096, SEQ_BLOCK = 1024, SEQ_BLOCKS = SEQ_CHUNK / SEQ_BLOCK;
__global__ void __launch_bounds__(256) seq_sum_kernel(const float *__restrict__ terms, int n, float *__restrict__ out, float divide_by,
                                                      const int *wait_flags, int wait_n, int wait_skip, int wait_value, int *wait_err,
                                                      long long wait_limit) {
    __shared__ __align__(16) float buf[2][SEQ_CHUNK];
    __shared__ int s_part[8], s_ok[8];
    __shared__ float s_S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (wait_flags) {                                           // row-partitioned runs: every rank's terms must have landed
        if (tid < wait_n && tid != wait_skip) {
            const long long t0 = clock64();
            for (;;) {
                int v;
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(wait_flags + tid) : "memory");
                if (v >= wait_value) break;
                if (clock64() - t0 > wait_limit) { *wait_err = 1; break; }
                __nanosleep(64);
            }
        }
        __syncthreads();
    }
    const int n_chunks = (n + SEQ_CHUNK - 1) / SEQ_CHUNK;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(terms) & 15) == 0;
    // terms beyond n read as +0, which a sum passes over unchanged
    auto fetch = [&](int c, float4 (&v)[4]) {                   // chunk c -> registers (all loads in flight together)
        const int base = c * SEQ_CHUNK;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = (tid + k * 256) * 4;
            if (vec_ok && base + i + 4 <= n) v[k] = ld_stream_f4(reinterpret_cast<const float4 *>(terms + base + i));
            else {
                v[k].x = base + i < n ? ld_stream_f32(terms + base + i) : 0.f;
                v[k].y = base + i + 1 < n ? ld_stream_f32(terms + base + i + 1) : 0.f;
                v[k].z = base + i + 2 < n ? ld_stream_f32(terms + base + i + 2) : 0.f;
                v[k].w = base + i + 3 < n ? ld_stream_f32(terms + base + i + 3) : 0.f;
            }
        }
    };
    auto stash = [&](int c, const float4 (&v)[4]) {
#pragma unroll
        for (int k = 0; k < 4; k++) reinterpret_cast<float4 *>(buf[c & 1])[tid + k * 256] = v[k];
    };
    float4 v[4];
    if (n_chunks > 0) { fetch(0, v); stash(0, v); }
    if (tid == 0) s_S = 0.f;
    __syncthreads();
    float S = 0.f;                                              // warp 0's copy is the truth; s_S broadcasts it once per chunk
    for (int c = 0; c < n_chunks; c++) {
        const float *src = buf[c & 1];
        if (c + 1 < n_chunks) fetch(c + 1, v);                  // in flight during the rounding below
        // speculative rounding: warp w takes rows [16 (w & 1), +16) of block w / 2
        const uint32_t sb0 = __float_as_uint(s_S);
        const int e0 = (int)((sb0 >> 23) & 0xff);
        const bool spec = !(sb0 >> 31) && e0 >= 30 && e0 <= 250;
        {
            bool ok = spec;
            int r = 0;
            if (spec) {
                float t[16];
                const float *blk = src + (warp >> 1) * SEQ_BLOCK + (warp & 1) * 512;
#pragma unroll
                for (int j = 0; j < 16; j++) t[j] = blk[j * 32 + lane];
                r = seq_round_terms(e0, t, 16, ok);
            }
            ok = __all_sync(FULL, ok);
            r = __reduce_add_sync(FULL, r);                     // < 2^20 * 512: exact
            if (lane == 0) { s_part[warp] = r; s_ok[warp] = ok; }
        }
        if (c + 1 < n_chunks) stash(c + 1, v);
        __syncthreads();
        if (warp == 0) {
            for (int b = 0; b < SEQ_BLOCKS; b++) {
                const int first = c * SEQ_CHUNK + b * SEQ_BLOCK;
                if (first >= n) break;
                const uint32_t sb = __float_as_uint(S);
                const uint32_t m = (sb & 0x7fffffu) | 0x800000u;
                const uint32_t R = (uint32_t)(s_part[2 * b] + s_part[2 * b + 1]);
                if (s_ok[2 * b] && s_ok[2 * b + 1] && !(sb >> 31) && (int)((sb >> 23) & 0xff) == e0 && m + R < 0x1000000u) {
                    S = __uint_as_float((sb & 0x7f800000u) | ((m + R) & 0x7fffffu));
                    continue;
                }
                {                                               // the whole block again, in the binade S is in NOW (the speculation
                    float cur[SEQ_BLOCK / 32];                  // fails for every block behind a binade crossing in the same chunk)
#pragma unroll
                    for (int j = 0; j < SEQ_BLOCK / 32; j++) cur[j] = src[b * SEQ_BLOCK + j * 32 + lane];
                    if (seq_fast_block(S, cur, SEQ_BLOCK / 32)) continue;
                }
#pragma unroll 1
                for (int j = 0; j < SEQ_BLOCK / 32; j++) {      // a row of 32 at a time: fast if possible, else the reference's own loop
                    const float one = src[b * SEQ_BLOCK + j * 32 + lane];
                    if (seq_fast_block(S, &one, 1)) continue;
                    for (int l = 0; l < 32; l++) S = S + src[b * SEQ_BLOCK + j * 32 + l];
                }
            }
            if (lane == 0) s_S = S;
        }
        __syncthreads();
    }
    if (tid == 0) { out[0] = divide_by != 0.f ? S / divide_by : S; }
}

}  // namespace

extern "C" {

int gcnk_relu_fw(float *x, uint32_t *mask_bits, int64_t n, int training, gcnk_stream_t stream) {
    GCNK_REQUIRE(x && n >= 0 && (!training || mask_bits), "bad arguments");
    if (!n) return GCNK_OK;
    relu_fw_kernel<<<ew_grid((n + 31) / 32), 256, 0, S(stream)>>>(x, mask_bits, n, training);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_relu_bw(float *grad, const uint32_t *mask_bits, int64_t n, gcnk_stream_t stream) {
    GCNK_REQUIRE(grad && mask_bits && n >= 0, "bad arguments");
    if (!n) return GCNK_OK;
    relu_bw_kernel<<<ew_grid(n), 256, 0, S(stream)>>>(grad, mask_bits, n);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_dropout_apply(float *x, const uint32_t *keep_bits, int64_t n, float p, gcnk_stream_t stream) {
    GCNK_REQUIRE(x && keep_bits && n >= 0, "bad arguments");
    if (!n) return GCNK_OK;
    const float scale = 1 / (1 - p);                                       // module.cpp:212
    dropout_apply_kernel<<<ew_grid(n), 256, 0, S(stream)>>>(x, keep_bits, n, scale);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_set_truth(int *truth, const int *split, const int *label, int current_split, int n, gcnk_stream_t stream) {
    GCNK_REQUIRE(truth && split && label && n >= 0, "bad arguments");
    if (!n) return GCNK_OK;
    set_truth_kernel<<<(n + 255) / 256, 256, 0, S(stream)>>>(truth, split, label, current_split, n);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

size_t gcnk_softmax_ce_workspace(int n, int c) {
    (void)c;
    return sizeof(CePartial) * (size_t)std::max(1, (n + 255) / 256);
}

int gcnk_softmax_ce(float *logits, const int *truth, float *grad, int n, int c, int training, gcnk_ce_result *d_result,
                    float *workspace, size_t workspace_bytes, gcnk_stream_t stream) {
    GCNK_REQUIRE(logits && truth && d_result && n >= 0 && c > 0 && (!training || grad), "bad arguments");
    GCNK_REQUIRE(workspace && workspace_bytes >= gcnk_softmax_ce_workspace(n, c), "workspace too small");
    cudaStream_t st = S(stream);
    const int parts = std::max(1, (n + 255) / 256);
    CePartial *partials = reinterpret_cast<CePartial *>(workspace);
    ce_rows_kernel<<<parts, 256, 0, st>>>(logits, truth, training ? grad : nullptr, n, c, training, partials);
    GCNK_LAUNCHED();
    ce_finish_kernel<<<1, 256, 0, st>>>(partials, parts, d_result);
    GCNK_LAUNCHED();
    if (training && n) {
        const int64_t total = (int64_t)n * c;
        ce_scale_grad_kernel<<<ew_grid(total), 256, 0, st>>>(grad, total, d_result);
        GCNK_LAUNCHED();
    }
    return GCNK_OK;
}

int gcnk_accuracy(const float *logits, const int *truth, int n, int c, int *d_wrong_total2, gcnk_stream_t stream) {
    GCNK_REQUIRE(logits && truth && d_wrong_total2 && n >= 0 && c > 0, "bad arguments");
    cudaStream_t st = S(stream);
    const int parts = std::max(1, (n + 255) / 256);
    static CePartial *scratch[64] = {nullptr};
    static int scratch_parts[64] = {0};
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    if (scratch_parts[dev] < parts) {
        GCNK_CUDA(cudaStreamSynchronize(st));
        if (scratch[dev]) GCNK_CUDA(cudaFree(scratch[dev]));
        GCNK_CUDA(cudaMalloc(&scratch[dev], sizeof(CePartial) * parts));
        scratch_parts[dev] = parts;
    }
    accuracy_kernel<<<parts, 256, 0, st>>>(logits, truth, n, c, scratch[dev]);
    GCNK_LAUNCHED();
    accuracy_finish_kernel<<<1, 256, 0, st>>>(scratch[dev], parts, d_wrong_total2);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_sequential_sum(const float *terms, int n, float *d_out, float divide_by, const int *d_wait_flags, int n_flags, int skip,
                        int wait_value, int *d_err, gcnk_stream_t stream) {
    GCNK_REQUIRE(terms && d_out && n >= 0 && (!d_wait_flags || (n_flags > 0 && n_flags <= 32 && d_err)), "bad arguments");
    prefer_carveout(seq_sum_kernel);
    seq_sum_kernel<<<1, 256, 0, S(stream)>>>(terms, n, d_out, divide_by, d_wait_flags, n_flags, skip, wait_value, d_err, peer_spin_cycles());
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_sum_squares(const float *w, int64_t n, float *d_out, gcnk_stream_t stream) {
    GCNK_REQUIRE(w && d_out && n >= 0, "bad arguments");
    sum_squares_kernel<<<1, 1024, 0, S(stream)>>>(w, n, d_out);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_adam_step(const gcnk_adam_tensor *h_tensors, int count, float step_size, float beta1, float beta2, float eps,
                   float weight_decay, float *d_sumsq, gcnk_stream_t stream) {
    GCNK_REQUIRE(h_tensors && count > 0 && count <= ADAM_MAX, "1..8 tensors per call");
    AdamPack pack = {};
    int max_size = 0;
    for (int i = 0; i < count; i++) {
        GCNK_REQUIRE(h_tensors[i].data && h_tensors[i].grad && h_tensors[i].m && h_tensors[i].v && h_tensors[i].size >= 0, "bad tensor");
        pack.t[i] = h_tensors[i];
        max_size = std::max(max_size, h_tensors[i].size);
    }
    if (max_size) {
        dim3 grid(std::max(1, std::min((max_size + 255) / 256, sm_count() * 4)), count, 1);
        adam_kernel<<<grid, 256, 0, S(stream)>>>(pack, step_size, beta1, beta2, eps, weight_decay);
        GCNK_LAUNCHED();
    }
    if (d_sumsq) return gcnk_sum_squares(h_tensors[0].data, h_tensors[0].size, d_sumsq, stream);
    return GCNK_OK;
}

}  // extern "C"
