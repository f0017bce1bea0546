// wide.cu — the row-local elementwise pieces of the WIDE-hidden plan (ogbn-products shape: 100 features -> hidden 256 ->
// 47 classes), where hidden*classes no longer fits the fused layer-2 kernel (layer2.cu) and the contractions are real
// GEMMs (matmul_tc.cu).  Reference semantics: Dropout (module.cpp:207-233), ReLU (module.cpp:175-194),
// CrossEntropyLoss (module.cpp:124-161), get_accuracy (gcn.cpp:83-96).
//
// The wide plan re-orders layer 1 to (A_hat * drop(X)) * W1 — the gather runs at the input width (100) instead of the
// hidden width (256) and its result is reused for the W1 gradient, dW1 = (A_hat drop(X))^T dZ1, so NO gather runs at
// width 256 in either direction — and keeps layer 2 in the reference's order A_hat * (H1 * W2) at width 47 (stored
// with a zero 48th column so that every row is 16-byte aligned: TMA-addressable and float4-gatherable).
#include <algorithm>

#include "common.cuh"

using namespace gcnk;

namespace {

// out[i, :] = dinv[i] * (keep bit(i*f + j) ? x[i, j] * scale : 0)     (keep == NULL: keep everything, scale ignored)
// one thread per 4 consecutive elements of a row (f % 4 == 0), bits fetched as the 4-bit nibble of the flat stream
__global__ void __launch_bounds__(256) drop_scale_rows_kernel(const float4 *__restrict__ x, const uint32_t *__restrict__ keep, const float *__restrict__ dinv,
                                                              float4 *__restrict__ out, int64_t n_vec, int vec_per_row, float scale) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_vec; i += stride) {
        const float4 v = ld_stream_f4(x + i);
        const float d = dinv[i / vec_per_row];
        float4 o;
        if (keep) {
            const int64_t b = i * 4;                                   // flat bit index of v.x; 4 | b so the nibble never straddles a word
            const uint32_t k = (keep[b >> 5] >> (b & 31)) & 0xfu;
            const float ds = d * scale;
            o.x = (k & 1u) ? v.x * ds : 0.f; o.y = (k & 2u) ? v.y * ds : 0.f; o.z = (k & 4u) ? v.z * ds : 0.f; o.w = (k & 8u) ? v.w * ds : 0.f;
        } else {
            o = make_float4(v.x * d, v.y * d, v.z * d, v.w * d);
        }
        out[i] = o;
    }
}

// in place: h = (z > 0 && keep) ? z * scale : 0; mask bit = (z > 0) && keep.  One float4 per thread (coalesced); the eight
// lanes that share a 32-element mask word combine their nibbles with three shuffles.  n % 4 == 0.
__global__ void __launch_bounds__(256) relu_dropout_fw_kernel(float4 *__restrict__ z, const uint32_t *__restrict__ keep, uint32_t *__restrict__ mask,
                                                              int64_t n_vec, float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < n_vec; base += stride) {
        const int64_t i = base + lane;
        uint32_t nib = 0;
        if (i < n_vec) {
            float4 v = z[i];
            const int64_t bit = i * 4;
            const uint32_t k = keep ? (keep[bit >> 5] >> (bit & 31)) & 0xfu : 0xfu;
            const bool a = v.x > 0.f && (k & 1u), b = v.y > 0.f && (k & 2u), c = v.z > 0.f && (k & 4u), d = v.w > 0.f && (k & 8u);
            v.x = a ? v.x * scale : 0.f; v.y = b ? v.y * scale : 0.f; v.z = c ? v.z * scale : 0.f; v.w = d ? v.w * scale : 0.f;
            z[i] = v;
            nib = (uint32_t)a | (uint32_t)b << 1 | (uint32_t)c << 2 | (uint32_t)d << 3;
        }
        uint32_t word = nib << (4 * (lane & 7));
        word |= __shfl_xor_sync(FULL, word, 1);
        word |= __shfl_xor_sync(FULL, word, 2);
        word |= __shfl_xor_sync(FULL, word, 4);
        if (mask && (lane & 7) == 0 && i < n_vec) mask[i >> 3] = word;
    }
}

// in place: g = mask bit ? g * scale : 0     (Dropout backward then ReLU backward, module.cpp:186-194,226-233)
__global__ void __launch_bounds__(256) mask_scale_bw_kernel(float4 *__restrict__ g, const uint32_t *__restrict__ mask, int64_t n_vec, float scale) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_vec; i += stride) {
        float4 v = g[i];
        const int64_t bit = i * 4;
        const uint32_t k = (mask[bit >> 5] >> (bit & 31)) & 0xfu;
        v.x = (k & 1u) ? v.x * scale : 0.f; v.y = (k & 2u) ? v.y * scale : 0.f; v.z = (k & 4u) ? v.z * scale : 0.f; v.w = (k & 8u) ? v.w * scale : 0.f;
        g[i] = v;
    }
}

// dst[r, 0..ld) = {src[r, 0..c), 0...}  /  dst[r, 0..c) = src[r, 0..c) of a pitch-ld source
__global__ void pad_cols_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int c, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld) return;
    const int r = i / ld, j = i % ld;
    dst[i] = j < c ? src[(size_t)r * c + j] : 0.f;
}
__global__ void unpad_cols_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int c, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * c) return;
    const int r = i / c, j = i % c;
    dst[i] = src[(size_t)r * ld + j];
}

struct CePart { float loss; int count; int wrong; int pad; };

// Softmax cross-entropy + accuracy over rows of pitch ld (c <= 128 classes), one warp per labelled row, rows taken in
// blocks of 32 so that split/label are read coalesced and unlabelled rows cost nothing but their (zero) gradient row.
// Training: grad_scaled[s, :] = dinv[s] * (softmax - onehot) / count  (pitch ld, padding columns 0), zero rows for
// unlabelled nodes: the pre-scaled source of the backward GraphSum.
__global__ void __launch_bounds__(256) ce_rows_ld_kernel(const float *__restrict__ logits, int ld, const int *__restrict__ split, const int *__restrict__ label,
                                                         int current_split, int n, int c, int training, float count_f, const float *__restrict__ dinv,
                                                         float *__restrict__ grad_scaled, CePart *__restrict__ partials, float *__restrict__ terms,
                                                         const int *__restrict__ term_index) {
    constexpr int CPL = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    float loss = 0.f;
    int count = 0, wrong = 0;
    const int n_blocks = (n + 31) / 32, total_warps = gridDim.x * warps;
    for (int blk = blockIdx.x * warps + warp; blk < n_blocks; blk += total_warps) {
        const int base = blk * 32, my_row = base + lane;
        const int my_truth = (my_row < n && split[my_row] == current_split) ? label[my_row] : -1;   // set_truth (gcn.cpp:78-81)
        unsigned todo = __ballot_sync(FULL, my_truth >= 0);
        if (training) {
            const unsigned labelled = todo;
            for (int r = 0; r < 32 && base + r < n; r++)
                if (!((labelled >> r) & 1u))
                    for (int j = lane; j < ld; j += 32) grad_scaled[(size_t)(base + r) * ld + j] = 0.f;
        }
        if (terms && !term_index && my_row < n && my_truth < 0) terms[my_row] = 0.f;
        while (todo) {
            const int r_in = __ffs(todo) - 1;
            todo &= todo - 1;
            const int s = base + r_in;
            const int truth = __shfl_sync(FULL, my_truth, r_in);
            const float *row = logits + (size_t)s * ld;
            float lg[CPL], mx = -1e30f;
#pragma unroll
            for (int t = 0; t < CPL; t++) {
                const int cls = lane + 32 * t;
                lg[t] = cls < c ? row[cls] : 0.f;
                if (cls < c) mx = fmaxf(mx, lg[t]);
            }
            mx = warp_max(mx);
            float ex[CPL], sum = 0.f;
#pragma unroll
            for (int t = 0; t < CPL; t++) {
                ex[t] = (lane + 32 * t < c) ? expf(lg[t] - mx) : 0.f;
                sum += ex[t];
            }
            sum = warp_sum(sum);
            float tl = 0.f;
#pragma unroll
            for (int t = 0; t < CPL; t++) if (t == truth / 32) tl = lg[t];
            tl = __shfl_sync(FULL, tl, truth % 32);
            bool wr = false;
#pragma unroll
            for (int t = 0; t < CPL; t++) wr |= (lane + 32 * t < c) && lg[t] > tl;      // strict: ties count as correct (gcn.cpp:88-93)
            wr = __any_sync(FULL, wr);
            count++;
            wrong += wr;
            const float term = logf(sum) - (tl - mx);
            loss += term;
            if (terms && lane == 0) terms[term_index ? term_index[s] : s] = term;
            if (training) {
                const float di = dinv[s];
#pragma unroll
                for (int t = 0; t < CPL; t++) {
                    const int cls = lane + 32 * t;
                    if (cls < ld) {
                        float g = 0.f;
                        if (cls < c) {
                            g = ex[t] / sum;
                            if (cls == truth) g -= 1.0f;
                            g = g / count_f;                                             // grad /= count (module.cpp:156-158)
                        }
                        grad_scaled[(size_t)s * ld + cls] = di * g;
                    }
                }
            }
        }
    }
    __shared__ float s_loss[8];
    __shared__ int s_count[8], s_wrong[8];
    if (lane == 0) { s_loss[warp] = loss; s_count[warp] = count; s_wrong[warp] = wrong; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f; int cn = 0, w2 = 0;
        for (int w = 0; w < warps; w++) { l += s_loss[w]; cn += s_count[w]; w2 += s_wrong[w]; }
        partials[blockIdx.x] = CePart{l, cn, w2, 0};
    }
}

__global__ void __launch_bounds__(256) ce_ld_finish_kernel(const CePart *__restrict__ partials, int parts, gcnk_ce_result *__restrict__ result,
                                                           float *__restrict__ red4) {
    __shared__ float s_loss[256];
    __shared__ int s_count[256], s_wrong[256];
    float l = 0.f; int cn = 0, wr = 0;
    for (int b = threadIdx.x; b < parts; b += 256) { l += partials[b].loss; cn += partials[b].count; wr += partials[b].wrong; }
    s_loss[threadIdx.x] = l; s_count[threadIdx.x] = cn; s_wrong[threadIdx.x] = wr;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_loss[threadIdx.x] += s_loss[threadIdx.x + o]; s_count[threadIdx.x] += s_count[threadIdx.x + o]; s_wrong[threadIdx.x] += s_wrong[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        result->loss = s_loss[0] / (float)s_count[0];
        result->count = s_count[0]; result->wrong = s_wrong[0]; result->pad = 0;
        red4[0] = s_loss[0]; red4[1] = (float)s_count[0]; red4[2] = (float)s_wrong[0]; red4[3] = 0.f;
    }
}

int ce_grid(int n) { return std::max(1, std::min(((n + 31) / 32 + 7) / 8, sm_count() * 8)); }

}  // namespace

extern "C" {

int gcnk_drop_scale_rows(const float *x, int rows, int f, const uint32_t *keep_bits, float scale, const float *d_dinv, float *out,
                         gcnk_stream_t stream) {
    GCNK_REQUIRE(x && out && d_dinv && rows >= 0 && f > 0 && f % 4 == 0, "bad arguments (f must be a multiple of 4)");
    GCNK_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0, "x and out must be 16-byte aligned");
    const int64_t n_vec = (int64_t)rows * (f / 4);
    if (!n_vec) return GCNK_OK;
    const int grid = (int)std::min<int64_t>((n_vec + 255) / 256, (int64_t)sm_count() * 16);
    drop_scale_rows_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float4 *>(x), keep_bits, d_dinv, reinterpret_cast<float4 *>(out), n_vec,
                                                        f / 4, scale);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_relu_dropout_fw(float *z, int64_t n, const uint32_t *keep_bits, float scale, uint32_t *mask_bits, gcnk_stream_t stream) {
    GCNK_REQUIRE(z && n >= 0 && n % 4 == 0 && reinterpret_cast<uintptr_t>(z) % 16 == 0, "bad arguments (n must be a multiple of 4)");
    if (!n) return GCNK_OK;
    const int64_t n_vec = n / 4;
    const int grid = (int)std::min<int64_t>((n_vec + 255) / 256, (int64_t)sm_count() * 16);
    relu_dropout_fw_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<float4 *>(z), keep_bits, mask_bits, n_vec, scale);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_mask_scale_bw(float *g, int64_t n, const uint32_t *mask_bits, float scale, gcnk_stream_t stream) {
    GCNK_REQUIRE(g && mask_bits && n >= 0 && n % 4 == 0 && reinterpret_cast<uintptr_t>(g) % 16 == 0, "bad arguments (n must be a multiple of 4)");
    if (!n) return GCNK_OK;
    const int64_t n_vec = n / 4;
    const int grid = (int)std::min<int64_t>((n_vec + 255) / 256, (int64_t)sm_count() * 16);
    mask_scale_bw_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<float4 *>(g), mask_bits, n_vec, scale);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_pad_cols(const float *src, float *dst, int rows, int c, int ld, gcnk_stream_t stream) {
    GCNK_REQUIRE(src && dst && rows >= 0 && c > 0 && ld >= c, "bad arguments");
    if (!rows) return GCNK_OK;
    pad_cols_kernel<<<(rows * ld + 255) / 256, 256, 0, S(stream)>>>(src, dst, rows, c, ld);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_unpad_cols(const float *src, float *dst, int rows, int c, int ld, gcnk_stream_t stream) {
    GCNK_REQUIRE(src && dst && rows >= 0 && c > 0 && ld >= c, "bad arguments");
    if (!rows) return GCNK_OK;
    unpad_cols_kernel<<<(rows * c + 255) / 256, 256, 0, S(stream)>>>(src, dst, rows, c, ld);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

size_t gcnk_ce_rows_workspace(int n) { return 16 + (size_t)ce_grid(n) * sizeof(CePart); }

int gcnk_ce_rows(const float *logits, int ld, const int *split, const int *label, int current_split, int n, int c, int training, int count,
                 const float *d_dinv, float *grad_scaled, gcnk_ce_result *d_result, float *workspace, size_t workspace_bytes, float *loss_terms,
                 const int *term_index, gcnk_stream_t stream) {
    GCNK_REQUIRE(logits && split && label && d_result && workspace && n >= 0 && c > 0 && c <= 128 && ld >= c && ld <= 128, "bad arguments");
    GCNK_REQUIRE(!training || (grad_scaled && d_dinv), "training needs the gradient buffer and dinv");
    GCNK_REQUIRE(workspace_bytes >= gcnk_ce_rows_workspace(n), "workspace too small");
    const int grid = ce_grid(n);
    CePart *parts = reinterpret_cast<CePart *>(workspace + 4);
    ce_rows_ld_kernel<<<grid, 256, 0, S(stream)>>>(logits, ld, split, label, current_split, n, c, training, (float)count, d_dinv, grad_scaled, parts,
                                                   loss_terms, term_index);
    GCNK_LAUNCHED();
    ce_ld_finish_kernel<<<1, 256, 0, S(stream)>>>(parts, grid, d_result, workspace);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // extern "C"
