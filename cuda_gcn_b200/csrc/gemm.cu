// gemm.cu — the three dense products of a wide-hidden GCN layer with explicit row pitches (include/gcnk.h:
// gcnk_matmul_nn / _nt / _tn), i.e. Matmul::forward, the dA and the dB halves of Matmul::backward (reference
// src/seq/module.cpp:11-42; GPU kernels src/cuda/cuda_kernel.cu:6-96) for operands that live inside padded buffers:
//
//   nn   C[m x n] = A[m x k] * B[k x n]            (+ optional row scale: the d^-1/2 pre-scale of the next GraphSum)
//   nt   C[m x n] = A[m x k] * B[n x k]^T          (dA = dC * B^T)
//   tn   C[ka x n] = A[m x ka]^T * B[m x n]        (dB = A^T * dC; m, the node dimension, is the contraction)
//
// Dispatch: the tcgen05 kernels of matmul_tc.cu when the shape qualifies (node dimension >= 1024, 16-byte-aligned
// pitches, operands that fit the shared-memory plan) — 3xTF32 split products, fp32-accurate — otherwise the register-
// tiled fp32 SIMT kernel below (any shape).  GCNK_NO_TCGEN05=1 forces the SIMT path (parity tests run both).
#include <algorithm>

#include "common.cuh"

using namespace gcnk;

namespace gcnk {   // matmul_tc.cu
bool matmul_tc_nn_supported(int m, int k, int n, int lda, int ldc, bool b_is_nk);
int matmul_tc_nn(const float *a, int lda, const float *b, int ldb, bool b_is_nk, float *c, int ldc, int m, int k, int n, const float *row_scale,
                 cudaStream_t st, const uint32_t *keep = nullptr, float out_scale = 1.0f, int relu = 0);
bool matmul_tc_tn_supported(int m, int ka, int n, int lda, int ldb);
size_t matmul_tc_tn_workspace(int m, int ka, int n);
int matmul_tc_tn(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int ka, int n, float *ws, size_t ws_bytes, cudaStream_t st,
                 const uint32_t *keep = nullptr, float out_scale = 1.0f);
}

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

// C[M x N] (pitch ldc) = op(A) * op(B) over k in [k_lo, k_hi);  TA: A is stored [K x M] (pitch lda) and used transposed,
// else [M x K];  TB: B is stored [N x K] (pitch ldb), else [K x N].  blockIdx.z = split-K part, written at C + z*M*ldc.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_ld_kernel(const float *__restrict__ A, int lda, const float *__restrict__ B, int ldb, float *__restrict__ C,
                                                      int ldc, int M, int N, int K, int k_chunk, const float *__restrict__ row_scale) {
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_lo = blockIdx.z * k_chunk, k_hi = min(K, k_lo + k_chunk);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;
    for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
        for (int i = threadIdx.x; i < BK * BM; i += 256) {
            int kk, mm;
            if (TA) { mm = i % BM; kk = i / BM; } else { kk = i % BK; mm = i / BK; }
            const int gm = m0 + mm, gk = k0 + kk;
            float v = 0.f;
            if (gm < M && gk < k_hi) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[kk][mm] = v;
        }
        for (int i = threadIdx.x; i < BK * BN; i += 256) {
            int kk, nn;
            if (TB) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
            const int gn = n0 + nn, gk = k0 + kk;
            float v = 0.f;
            if (gn < N && gk < k_hi) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *out = C + (size_t)blockIdx.z * M * ldc;
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gm = m0 + ty + 16 * i;
        if (gm >= M) continue;
        const float rs = row_scale ? row_scale[gm] : 1.0f;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int gn = n0 + tx + 16 * j;
            if (gn < N) out[(size_t)gm * ldc + gn] = row_scale ? rs * acc[i][j] : acc[i][j];
        }
    }
}

__global__ void reduce_parts_ld_kernel(const float *__restrict__ partials, float *__restrict__ out, int rows, int cols, int ld_part, int ldc, int parts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i % cols;
    float s = 0.f;
    for (int b = 0; b < parts; b++) s += partials[((size_t)b * rows + r) * ld_part + c];   // part order: fixed => deterministic
    out[(size_t)r * ldc + c] = s;
}

int simt_parts(int m) { return std::max(1, std::min(sm_count() * 2, (m + 511) / 512)); }

bool tc_off() {
    static const bool off = getenv("GCNK_NO_TCGEN05") && atoi(getenv("GCNK_NO_TCGEN05")) != 0;
    return off;
}

}  // namespace

extern "C" {

int gcnk_matmul_nn(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int k, int n, const float *row_scale,
                   gcnk_stream_t stream) {
    GCNK_REQUIRE(a && b && c && m >= 0 && k > 0 && n > 0 && lda >= k && ldb >= n && ldc >= n, "bad arguments");
    if (m == 0) return GCNK_OK;
    if (!tc_off() && matmul_tc_nn_supported(m, k, n, lda, ldc, false) && reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(c) % 16 == 0) {
        const int rc = matmul_tc_nn(a, lda, b, ldb, false, c, ldc, m, k, n, row_scale, S(stream));
        if (rc != GCNK_EUNSUPPORTED) return rc;
    }
    dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM, 1);
    gemm_ld_kernel<false, false><<<grid, 256, 0, S(stream)>>>(a, lda, b, ldb, c, ldc, m, n, k, k, row_scale);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_matmul_nt(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int k, int n, gcnk_stream_t stream) {
    GCNK_REQUIRE(a && b && c && m >= 0 && k > 0 && n > 0 && lda >= k && ldb >= k && ldc >= n, "bad arguments");
    if (m == 0) return GCNK_OK;
    if (!tc_off() && matmul_tc_nn_supported(m, k, n, lda, ldc, true) && reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(c) % 16 == 0) {
        const int rc = matmul_tc_nn(a, lda, b, ldb, true, c, ldc, m, k, n, nullptr, S(stream));
        if (rc != GCNK_EUNSUPPORTED) return rc;
    }
    dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM, 1);
    gemm_ld_kernel<false, true><<<grid, 256, 0, S(stream)>>>(a, lda, b, ldb, c, ldc, m, n, k, k, nullptr);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

// Dropout-on-read forms for the dense feature transform (SparseMatmul with a dense X, module.cpp:47-77, with the Dropout of
// module.cpp:207-224 applied while the tile is split in shared memory): tensor-core path only.
int gcnk_dense_transform_tc(const float *xp, int ld, int m, int n, const float *w, float *c, int p, const uint32_t *drop_bits, float drop_scale,
                            const float *row_scale, int relu, gcnk_stream_t stream) {
    GCNK_REQUIRE(xp && w && c && m >= 0 && n > 0 && p > 0 && ld >= n, "bad arguments");
    if (m == 0) return GCNK_OK;
    if (tc_off() || !matmul_tc_nn_supported(m, n, p, ld, p, false) || reinterpret_cast<uintptr_t>(xp) % 16) {
        set_error("gcnk_dense_transform_tc: shape not supported by the tcgen05 kernel");
        return GCNK_EUNSUPPORTED;
    }
    return matmul_tc_nn(xp, ld, w, p, false, c, p, m, n, p, row_scale, S(stream), drop_bits, drop_bits ? drop_scale : 1.0f, relu);
}

size_t gcnk_dense_transform_bw_tc_workspace(int m, int n, int p) { return matmul_tc_tn_workspace(m, n, p); }

int gcnk_dense_transform_bw_tc(const float *xp, int ld, int m, int n, const float *g, float *w_grad, int p, const uint32_t *drop_bits, float drop_scale,
                               float *workspace, size_t workspace_bytes, gcnk_stream_t stream) {
    GCNK_REQUIRE(xp && g && w_grad && m >= 0 && n > 0 && p > 0 && ld >= n, "bad arguments");
    if (tc_off() || !matmul_tc_tn_supported(m, n, p, ld, p) || reinterpret_cast<uintptr_t>(xp) % 16 || reinterpret_cast<uintptr_t>(g) % 16) {
        set_error("gcnk_dense_transform_bw_tc: shape not supported by the tcgen05 kernel");
        return GCNK_EUNSUPPORTED;
    }
    return matmul_tc_tn(xp, ld, g, p, w_grad, p, m, n, p, workspace, workspace_bytes, S(stream), drop_bits, drop_bits ? drop_scale : 1.0f);
}

size_t gcnk_matmul_tn_workspace(int m, int ka, int n) {
    const size_t simt = simt_parts(m) > 1 ? sizeof(float) * (size_t)simt_parts(m) * ka * n : 0;
    return std::max(simt, matmul_tc_tn_workspace(m, ka, n));
}

int gcnk_matmul_tn(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int ka, int n, float *workspace, size_t workspace_bytes,
                   gcnk_stream_t stream) {
    GCNK_REQUIRE(a && b && c && m >= 0 && ka > 0 && n > 0 && lda >= ka && ldb >= n && ldc >= n, "bad arguments");
    cudaStream_t st = S(stream);
    if (m == 0) {
        for (int r = 0; r < ka; r++) GCNK_CUDA(cudaMemsetAsync(c + (size_t)r * ldc, 0, sizeof(float) * n, st));
        return GCNK_OK;
    }
    GCNK_REQUIRE(workspace_bytes >= gcnk_matmul_tn_workspace(m, ka, n) && (workspace || !workspace_bytes), "workspace too small");
    if (!tc_off() && matmul_tc_tn_supported(m, ka, n, lda, ldb) && reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0) {
        const int rc = matmul_tc_tn(a, lda, b, ldb, c, ldc, m, ka, n, workspace, workspace_bytes, st);
        if (rc != GCNK_EUNSUPPORTED) return rc;
    }
    int parts = simt_parts(m);
    int k_chunk = (m + parts - 1) / parts;
    k_chunk = (k_chunk + BK - 1) / BK * BK;
    parts = (m + k_chunk - 1) / k_chunk;
    dim3 grid((n + BN - 1) / BN, (ka + BM - 1) / BM, parts);
    if (parts > 1) {
        gemm_ld_kernel<true, false><<<grid, 256, 0, st>>>(a, lda, b, ldb, workspace, n, ka, n, m, k_chunk, nullptr);   // parts of [ka x n], pitch n
        GCNK_LAUNCHED();
        reduce_parts_ld_kernel<<<(ka * n + 255) / 256, 256, 0, st>>>(workspace, c, ka, n, n, ldc, parts);
    } else {
        gemm_ld_kernel<true, false><<<grid, 256, 0, st>>>(a, lda, b, ldb, c, ldc, ka, n, m, k_chunk, nullptr);
    }
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // extern "C"
