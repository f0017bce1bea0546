// dense.cu — Matmul forward / backward-A / backward-B.  Replaces cuda_Matmul_forward_kernel,
// cuda_Matmul_backward_A_kernel and cuda_Matmul_backward_B_kernel (reference
// src/cuda/cuda_kernel.cu:6-96; CPU semantics src/seq/module.cpp:11-42).
//
// One register-tiled fp32 SIMT GEMM, C[M x N] = op(A)[M x K] * op(B)[K x N], 64x64 CTA tile, 4x4 per
// thread, with optional split-K: the reference's backward-B launches 2 CTAs that each loop over all
// 7,281 row tiles serially; here the node dimension is split across CTAs and the per-CTA
// partials are summed in CTA order (deterministic).  These are the layer-2 products at hidden 16 /
// 41 classes (AI ~ 10 FLOP/B, HBM-bound); in the fused plan they are absorbed into gcnk_layer2_fused
// and never launched.  Wide hidden layers (products shape, hidden 256) are the tcgen05 follow-up.
#include <algorithm>

#include "common.cuh"

using namespace gcnk;

namespace gcnk {   // matmul_tc.cu: tcgen05 / TMEM / TMA path for wide reductions
bool matmul_tc_supported(int m, int k, int n);
int matmul_tc_fw(const float *a, const float *b, float *c, int m, int k, int n, cudaStream_t st);
}

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

// TA: A is stored [K x M] (used transposed); TB: B is stored [N x K].
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_kernel(const float *__restrict__ A, const float *__restrict__ B,
                                                    float *__restrict__ C, int M, int N, int K, int k_chunk) {
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
        // stage the two tiles; the faster-varying thread index follows the contiguous storage dimension
        for (int i = threadIdx.x; i < BK * BM; i += 256) {
            int kk, mm;
            if (TA) { mm = i % BM; kk = i / BM; } else { kk = i % BK; mm = i / BK; }
            const int gm = m0 + mm, gk = k0 + kk;
            float v = 0.f;
            if (gm < M && gk < k_hi) v = TA ? A[(size_t)gk * M + gm] : A[(size_t)gm * K + gk];
            As[kk][mm] = v;
        }
        for (int i = threadIdx.x; i < BK * BN; i += 256) {
            int kk, nn;
            if (TB) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
            const int gn = n0 + nn, gk = k0 + kk;
            float v = 0.f;
            if (gn < N && gk < k_hi) v = TB ? B[(size_t)gn * K + gk] : B[(size_t)gk * N + gn];
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
    float *out = C + (size_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gm = m0 + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int gn = n0 + tx + 16 * j;
            if (gm < M && gn < N) out[(size_t)gm * N + gn] = acc[i][j];
        }
    }
}

__global__ void reduce_splitk_kernel(const float *__restrict__ partials, float *__restrict__ out, int elems, int parts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= elems) return;
    float s = 0.f;
    for (int b = 0; b < parts; b++) s += partials[(size_t)b * elems + i];
    out[i] = s;
}

int splitk_parts(int m) { return std::max(1, std::min(sm_count() * 2, (m + 255) / 256)); }

}  // namespace

extern "C" {

int gcnk_matmul_fw(const float *a, const float *b, float *c, int m, int n, int p, gcnk_stream_t stream) {
    GCNK_REQUIRE(a && b && c && m >= 0 && n > 0 && p > 0, "bad arguments");
    if (m == 0) return GCNK_OK;
    if (matmul_tc_supported(m, n, p) && reinterpret_cast<uintptr_t>(a) % 16 == 0) {
        const int rc = matmul_tc_fw(a, b, c, m, n, p, S(stream));
        if (rc != GCNK_EUNSUPPORTED) return rc;
    }
    dim3 grid((p + BN - 1) / BN, (m + BM - 1) / BM, 1);
    gemm_kernel<false, false><<<grid, 256, 0, S(stream)>>>(a, b, c, m, p, n, n);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_matmul_bw_a(const float *c_grad, const float *b, float *a_grad, int m, int n, int p, gcnk_stream_t stream) {
    GCNK_REQUIRE(c_grad && b && a_grad && m >= 0 && n > 0 && p > 0, "bad arguments");
    if (m == 0) return GCNK_OK;
    dim3 grid((n + BN - 1) / BN, (m + BM - 1) / BM, 1);
    gemm_kernel<false, true><<<grid, 256, 0, S(stream)>>>(c_grad, b, a_grad, m, n, p, p);   // [m x p] * (b[n x p])^T
    GCNK_LAUNCHED();
    return GCNK_OK;
}

size_t gcnk_matmul_bw_b_workspace(int m, int n, int p) {
    const int parts = splitk_parts(m);
    return parts > 1 ? sizeof(float) * (size_t)parts * n * p : 0;
}

int gcnk_matmul_bw_b(const float *a, const float *c_grad, float *b_grad, int m, int n, int p, float *workspace,
                     size_t workspace_bytes, gcnk_stream_t stream) {
    GCNK_REQUIRE(a && c_grad && b_grad && m >= 0 && n > 0 && p > 0, "bad arguments");
    cudaStream_t st = S(stream);
    if (m == 0) { GCNK_CUDA(cudaMemsetAsync(b_grad, 0, sizeof(float) * (size_t)n * p, st)); return GCNK_OK; }
    int parts = splitk_parts(m);
    if (parts > 1) GCNK_REQUIRE(workspace && workspace_bytes >= gcnk_matmul_bw_b_workspace(m, n, p), "workspace too small");
    int k_chunk = (m + parts - 1) / parts;
    k_chunk = (k_chunk + BK - 1) / BK * BK;
    parts = (m + k_chunk - 1) / k_chunk;
    dim3 grid((p + BN - 1) / BN, (n + BM - 1) / BM, parts);
    gemm_kernel<true, false><<<grid, 256, 0, st>>>(a, c_grad, parts > 1 ? workspace : b_grad, n, p, m, k_chunk);   // (a[m x n])^T * [m x p]
    GCNK_LAUNCHED();
    if (parts > 1) {
        const int elems = n * p;
        reduce_splitk_kernel<<<(elems + 255) / 256, 256, 0, st>>>(workspace, b_grad, elems, parts);
        GCNK_LAUNCHED();
    }
    return GCNK_OK;
}

}  // extern "C"
