// feature.cu — SparseMatmul forward/backward: the feature transform X*W1 and its weight gradient.
// Replaces cuda_SparseMatmul_forward/backward_kernel (reference src/cuda/cuda_kernel.cu:100-122;
// CPU semantics src/seq/module.cpp:47-77).
//
// Two layouts behind one handle (gcnk_spmat):
//   * CSR (cora/citeseer/pubmed-like bag-of-words rows): forward is a row gather of W rows, a
//     sub-warp per output row; backward gathers through a CSC view built once at create time, so
//     every b_grad element is produced by exactly one thread in a fixed order (the reference's
//     backward does unsynchronised += from different blocks).
//   * dense (Reddit-like: each row stores all n columns): the 4-byte index per value is never read,
//     halving the HBM traffic; forward is a register-tiled tall-skinny product with W staged in
//     shared memory, backward a split-over-rows reduction with per-CTA partials reduced in CTA order.
// Dropout of the input features is applied ON READ from a keep-bit stream; the stored feature values
// are never modified, so the reference's per-pass host->device restore (cuda_gcn.cu:81-83) disappears.
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace gcnk;

namespace gcnk {   // feature_tc.cu: tensor-core (3xTF32) dense path at p == 16
bool dense_tc_supported(int n, int p);
int dense_fw16_tc(const float *x, const float *w, float *c, int m, int n, const uint32_t *bits, int64_t nnz, float scale,
                  const float *row_scale, int relu, cudaStream_t st);
size_t dense_bw16_tc_parts(int m);
int dense_bw16_tc(const float *x, const float *g, float *b_grad, float *partials, int m, int n, const uint32_t *bits, int64_t nnz,
                  float scale, cudaStream_t st);
}

struct gcnk_spmat {
    const int *indptr = nullptr, *indices = nullptr;
    int m = 0, n = 0;
    int64_t nnz = 0;
    int dense = 0;
    // CSC view (built lazily, CSR layouts only): column j owns entries [csc_ptr[j], csc_ptr[j+1])
    int *csc_ptr = nullptr, *csc_row = nullptr, *csc_pos = nullptr;
    float *partials = nullptr; size_t partial_elems = 0;   // dense backward per-CTA partial sums
};

namespace {

__device__ __forceinline__ float dropped(float v, const uint32_t *bits, int64_t pos, float scale) {
    if (!bits) return v;
    return ((bits[pos >> 5] >> (pos & 31)) & 1u) ? v * scale : 0.f;
}

// ---------------------------------------------------------------------------- generic forward ----
// LP lanes per output row (LP = min(32, pow2 >= p)); each lane owns columns k = q + LP*t, t < KT,
// per pass over the row; rows with p > LP*KT take several passes.
template <int LP, int KT>
__global__ void __launch_bounds__(256) spmm_fw_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                                       const float *__restrict__ values, const float *__restrict__ b,
                                                       float *__restrict__ c, int m, int n, int p,
                                                       const uint32_t *__restrict__ drop_bits, float drop_scale,
                                                       const float *__restrict__ row_scale) {
    const int rows_per_block = 256 / LP;
    const int q = threadIdx.x % LP;
    for (int i = blockIdx.x * rows_per_block + threadIdx.x / LP; i < m; i += gridDim.x * rows_per_block) {
        const int beg = indptr[i], end = indptr[i + 1];
        const float rs = row_scale ? row_scale[i] : 1.f;
        for (int k0 = 0; k0 < p; k0 += LP * KT) {
            float acc[KT];
#pragma unroll
            for (int t = 0; t < KT; t++) acc[t] = 0.f;
            for (int jj = beg; jj < end; jj++) {
                const int j = indices ? indices[jj] : jj - beg;
                const float x = dropped(values[jj], drop_bits, jj, drop_scale);
                const float *brow = b + (size_t)j * p + k0;
#pragma unroll
                for (int t = 0; t < KT; t++) {
                    const int k = q + LP * t;
                    if (k0 + k < p) acc[t] = fmaf(x, __ldg(brow + k), acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < KT; t++) {
                const int k = k0 + q + LP * t;
                if (k < p) c[(size_t)i * p + k] = rs * acc[t];
            }
        }
    }
}

// --------------------------------------------------------------------------- generic backward ----
// One sub-warp per feature column j; entries of the column arrive in ascending row order (stable
// counting sort), i.e. the reference's accumulation order (module.cpp:68-74).
template <int LP, int KT>
__global__ void __launch_bounds__(256) spmm_bw_kernel(const int *__restrict__ csc_ptr, const int *__restrict__ csc_row,
                                                       const int *__restrict__ csc_pos, const float *__restrict__ values,
                                                       const float *__restrict__ c_grad, float *__restrict__ b_grad,
                                                       int n, int p, const uint32_t *__restrict__ drop_bits, float drop_scale) {
    const int cols_per_block = 256 / LP;
    const int q = threadIdx.x % LP;
    for (int j = blockIdx.x * cols_per_block + threadIdx.x / LP; j < n; j += gridDim.x * cols_per_block) {
        const int beg = csc_ptr[j], end = csc_ptr[j + 1];
        for (int k0 = 0; k0 < p; k0 += LP * KT) {
            float acc[KT];
#pragma unroll
            for (int t = 0; t < KT; t++) acc[t] = 0.f;
            for (int e = beg; e < end; e++) {
                const int pos = csc_pos[e];
                const float x = dropped(values[pos], drop_bits, pos, drop_scale);
                const float *grow = c_grad + (size_t)csc_row[e] * p + k0;
#pragma unroll
                for (int t = 0; t < KT; t++) {
                    const int k = q + LP * t;
                    if (k0 + k < p) acc[t] = fmaf(__ldg(grow + k), x, acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < KT; t++) {
                const int k = k0 + q + LP * t;
                if (k < p) b_grad[(size_t)j * p + k] = acc[t];
            }
        }
    }
}

// the dense layout's implicit CSC: column j = rows 0..m-1 at positions i*n + j
template <int LP, int KT>
__global__ void __launch_bounds__(256) spmm_bw_dense_generic_kernel(const float *__restrict__ values,
                                                                     const float *__restrict__ c_grad,
                                                                     float *__restrict__ b_grad, int m, int n, int p,
                                                                     const uint32_t *__restrict__ drop_bits, float drop_scale) {
    const int cols_per_block = 256 / LP;
    const int q = threadIdx.x % LP;
    for (int j = blockIdx.x * cols_per_block + threadIdx.x / LP; j < n; j += gridDim.x * cols_per_block) {
        for (int k0 = 0; k0 < p; k0 += LP * KT) {
            float acc[KT];
#pragma unroll
            for (int t = 0; t < KT; t++) acc[t] = 0.f;
            for (int i = 0; i < m; i++) {
                const int64_t pos = (int64_t)i * n + j;
                const float x = dropped(values[pos], drop_bits, pos, drop_scale);
                const float *grow = c_grad + (size_t)i * p + k0;
#pragma unroll
                for (int t = 0; t < KT; t++) {
                    const int k = q + LP * t;
                    if (k0 + k < p) acc[t] = fmaf(__ldg(grow + k), x, acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < KT; t++) {
                const int k = k0 + q + LP * t;
                if (k < p) b_grad[(size_t)j * p + k] = acc[t];
            }
        }
    }
}

// ------------------------------------------------------------------ dense forward, p == 16 ----
// C[m x 16] = X[m x n] * W[n x 16].  A warp takes 4 rows at a time; lane l owns the feature columns
// f = l, l+32, ...: it streams X[r][f] for the 4 rows (coalesced, read once, L1-bypassing), reads
// W[f][0..15] from shared memory (row stride 20 floats => conflict-free 128-bit reads) and keeps
// 4x16 partial dot products in registers, so each W read feeds 4 rows.  The 64 partials are then
// summed across the warp with a transposing butterfly (62 shuffles instead of 320) that leaves lane l
// holding outputs (row l/8, columns 2*(l%8), +1): one coalesced 256-byte store per row quad.
constexpr int DF_ROWS = 4, DF_P = 16, DF_WSTRIDE = 20;

__global__ void __launch_bounds__(256) dense_fw16_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                          float *__restrict__ c, int m, int n,
                                                          const uint32_t *__restrict__ drop_bits, float drop_scale,
                                                          const float *__restrict__ row_scale) {
    extern __shared__ __align__(16) float sw[];   // [n][DF_WSTRIDE]
    for (int i = threadIdx.x; i < n * DF_P; i += blockDim.x) sw[(i / DF_P) * DF_WSTRIDE + (i % DF_P)] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int n_quads = (m + DF_ROWS - 1) / DF_ROWS;
    for (int quad = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); quad < n_quads; quad += warps_total) {
        const int r0 = quad * DF_ROWS;
        float acc[DF_ROWS * DF_P];
#pragma unroll
        for (int i = 0; i < DF_ROWS * DF_P; i++) acc[i] = 0.f;
        for (int f = lane; f < n; f += 32) {
            float xv[DF_ROWS];
#pragma unroll
            for (int r = 0; r < DF_ROWS; r++) {
                const int row = r0 + r;
                float v = 0.f;
                if (row < m) {
                    const int64_t pos = (int64_t)row * n + f;
                    v = dropped(ld_stream_f32(x + pos), drop_bits, pos, drop_scale);
                }
                xv[r] = v;
            }
            const float4 *wr = reinterpret_cast<const float4 *>(sw + f * DF_WSTRIDE);
            const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
            const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
            for (int r = 0; r < DF_ROWS; r++)
#pragma unroll
                for (int k = 0; k < DF_P; k++) acc[r * DF_P + k] = fmaf(xv[r], wv[k], acc[r * DF_P + k]);
        }
        // transposing butterfly: after the step with offset `off` a lane keeps the half selected by its
        // (lane & off) bit; five steps take 64 values per lane down to 2, fully summed over the warp.
#pragma unroll
        for (int half = 32, off = 16; off >= 1; half >>= 1, off >>= 1) {
            const bool upper = lane & off;
#pragma unroll
            for (int i = 0; i < half; i++) {
                const float send = upper ? acc[i] : acc[i + half];
                const float keep = upper ? acc[i + half] : acc[i];
                acc[i] = keep + __shfl_xor_sync(FULL, send, off);
            }
        }
        const int row = r0 + lane / 8;
        if (row < m) {
            const float rs = row_scale ? row_scale[row] : 1.f;
            *reinterpret_cast<float2 *>(c + (size_t)row * DF_P + (lane % 8) * 2) = make_float2(rs * acc[0], rs * acc[1]);
        }
    }
}

// ----------------------------------------------------------------- dense backward, p == 16 ----
// Wgrad[n x 16] = X^T[n x m] * G[m x 16].  CTA b owns a contiguous slab of rows; thread t owns the
// feature columns f = t, t+256, ... (<= DB_FPT of them) and keeps their 16 partial sums in registers.
// Per row the CTA reads the X row once (coalesced) and the 64-byte G row as a broadcast.  Per-CTA
// partials go to a workspace and are summed in CTA order by a second kernel (deterministic).
constexpr int DB_FPT = 4, DB_P = 16;

__global__ void __launch_bounds__(256) dense_bw16_kernel(const float *__restrict__ x, const float *__restrict__ g,
                                                          float *__restrict__ partials, int m, int n, int rows_per_cta,
                                                          const uint32_t *__restrict__ drop_bits, float drop_scale) {
    const int r_lo = blockIdx.x * rows_per_cta, r_hi = min(m, r_lo + rows_per_cta);
    float acc[DB_FPT][DB_P];
#pragma unroll
    for (int a = 0; a < DB_FPT; a++)
#pragma unroll
        for (int k = 0; k < DB_P; k++) acc[a][k] = 0.f;
    for (int row = r_lo; row < r_hi; row++) {
        const float4 *gr = reinterpret_cast<const float4 *>(g + (size_t)row * DB_P);
        const float4 g0 = __ldg(gr), g1 = __ldg(gr + 1), g2 = __ldg(gr + 2), g3 = __ldg(gr + 3);
        const float gv[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
#pragma unroll
        for (int a = 0; a < DB_FPT; a++) {
            const int f = threadIdx.x + 256 * a;
            if (f < n) {
                const int64_t pos = (int64_t)row * n + f;
                const float xv = dropped(ld_stream_f32(x + pos), drop_bits, pos, drop_scale);
#pragma unroll
                for (int k = 0; k < DB_P; k++) acc[a][k] = fmaf(gv[k], xv, acc[a][k]);
            }
        }
    }
    float *out = partials + (size_t)blockIdx.x * n * DB_P;
#pragma unroll
    for (int a = 0; a < DB_FPT; a++) {
        const int f = threadIdx.x + 256 * a;
        if (f < n) {
            float4 *o = reinterpret_cast<float4 *>(out + (size_t)f * DB_P);
            o[0] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
            o[1] = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
            o[2] = make_float4(acc[a][8], acc[a][9], acc[a][10], acc[a][11]);
            o[3] = make_float4(acc[a][12], acc[a][13], acc[a][14], acc[a][15]);
        }
    }
}

__global__ void reduce_partials_kernel(const float *__restrict__ partials, float *__restrict__ out, int elems, int parts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= elems) return;
    float s = 0.f;
    for (int b = 0; b < parts; b++) s += partials[(size_t)b * elems + i];
    out[i] = s;
}

__global__ void dense_check_kernel(const int *indptr, const int *indices, int m, int n, int *not_dense) {
    const int i = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (i >= m) return;
    const int beg = indptr[i], end = indptr[i + 1];
    if (end - beg != n) { if (lane == 0) atomicOr(not_dense, 1); return; }
    for (int f = lane; f < n; f += 32)
        if (indices[beg + f] != f) { atomicOr(not_dense, 1); return; }
}

template <int LP>
int launch_fw_generic(const gcnk_spmat *sp, const float *values, const float *b, float *c, int p, const uint32_t *drop_bits,
                      float drop_scale, const float *row_scale, cudaStream_t st) {
    const int rows_per_block = 256 / LP;
    const int grid = std::max(1, std::min((sp->m + rows_per_block - 1) / rows_per_block, sm_count() * 32));
    spmm_fw_kernel<LP, 4><<<grid, 256, 0, st>>>(sp->indptr, sp->dense ? nullptr : sp->indices, values, b, c, sp->m, sp->n, p,
                                                drop_bits, drop_scale, row_scale);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int build_csc(gcnk_spmat *sp, cudaStream_t st) {
    std::vector<int> indptr((size_t)sp->m + 1), indices((size_t)sp->nnz);
    GCNK_CUDA(cudaMemcpyAsync(indptr.data(), sp->indptr, sizeof(int) * indptr.size(), cudaMemcpyDeviceToHost, st));
    if (sp->nnz) GCNK_CUDA(cudaMemcpyAsync(indices.data(), sp->indices, sizeof(int) * indices.size(), cudaMemcpyDeviceToHost, st));
    GCNK_CUDA(cudaStreamSynchronize(st));
    std::vector<int> ptr((size_t)sp->n + 1, 0), row((size_t)sp->nnz), pos((size_t)sp->nnz);
    for (int64_t e = 0; e < sp->nnz; e++) {
        const int j = indices[e];
        if (j < 0 || j >= sp->n) { set_error("gcnk_spmat: column id %d out of range [0,%d)", j, sp->n); return GCNK_EINVAL; }
        ptr[j + 1]++;
    }
    for (int j = 0; j < sp->n; j++) ptr[j + 1] += ptr[j];
    std::vector<int> cur(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < sp->m; i++)
        for (int e = indptr[i]; e < indptr[i + 1]; e++) {
            const int slot = cur[indices[e]]++;
            row[slot] = i;
            pos[slot] = e;
        }
    GCNK_CUDA(cudaMalloc(&sp->csc_ptr, sizeof(int) * ptr.size()));
    GCNK_CUDA(cudaMalloc(&sp->csc_row, sizeof(int) * std::max<size_t>(row.size(), 1)));
    GCNK_CUDA(cudaMalloc(&sp->csc_pos, sizeof(int) * std::max<size_t>(pos.size(), 1)));
    GCNK_CUDA(cudaMemcpyAsync(sp->csc_ptr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice, st));
    if (sp->nnz) {
        GCNK_CUDA(cudaMemcpyAsync(sp->csc_row, row.data(), sizeof(int) * row.size(), cudaMemcpyHostToDevice, st));
        GCNK_CUDA(cudaMemcpyAsync(sp->csc_pos, pos.data(), sizeof(int) * pos.size(), cudaMemcpyHostToDevice, st));
    }
    GCNK_CUDA(cudaStreamSynchronize(st));
    return GCNK_OK;
}

}  // namespace

extern "C" {

int gcnk_spmat_create(gcnk_spmat **out, const int *d_indptr, const int *d_indices, int m, int n, int64_t nnz,
                      gcnk_stream_t stream) {
    GCNK_REQUIRE(out && d_indptr && (d_indices || nnz == 0) && m >= 0 && n > 0 && nnz >= 0, "bad arguments");
    GCNK_REQUIRE(nnz <= INT32_MAX, "nnz must fit int32 (as the reference's std::vector<int> indptr)");
    cudaStream_t st = S(stream);
    gcnk_spmat *sp = new gcnk_spmat;
    sp->indptr = d_indptr; sp->indices = d_indices; sp->m = m; sp->n = n; sp->nnz = nnz;
    if (nnz == (int64_t)m * n && m > 0) {
        int *d_flag = nullptr, flag = 0;
        GCNK_CUDA(cudaMalloc(&d_flag, sizeof(int)));
        GCNK_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
        dense_check_kernel<<<(m + 7) / 8, 256, 0, st>>>(d_indptr, d_indices, m, n, d_flag);
        GCNK_LAUNCHED();
        GCNK_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        GCNK_CUDA(cudaStreamSynchronize(st));
        GCNK_CUDA(cudaFree(d_flag));
        sp->dense = !flag;
    }
    *out = sp;
    return GCNK_OK;
}

int gcnk_spmat_destroy(gcnk_spmat *sp) {
    if (!sp) return GCNK_OK;
    cudaFree(sp->csc_ptr); cudaFree(sp->csc_row); cudaFree(sp->csc_pos); cudaFree(sp->partials);
    delete sp;
    return GCNK_OK;
}

int gcnk_spmat_is_dense(const gcnk_spmat *sp, int *is_dense) {
    GCNK_REQUIRE(sp && is_dense, "null");
    *is_dense = sp->dense;
    return GCNK_OK;
}

int gcnk_spmm_fw(const gcnk_spmat *sp, const float *values, const float *b, float *c, int p, const uint32_t *drop_bits,
                 float drop_scale, const float *row_scale, gcnk_stream_t stream) {
    GCNK_REQUIRE(sp && values && b && c && p > 0, "bad arguments");
    cudaStream_t st = S(stream);
    if (sp->m == 0) return GCNK_OK;
    if (sp->dense && dense_tc_supported(sp->n, p) && reinterpret_cast<uintptr_t>(c) % 8 == 0 &&
        reinterpret_cast<uintptr_t>(values) % 8 == 0) {
        const int rc = dense_fw16_tc(values, b, c, sp->m, sp->n, drop_bits, sp->nnz, drop_scale, row_scale, 0, st);
        if (rc != GCNK_EUNSUPPORTED) return rc;
    }
    const size_t smem = sizeof(float) * (size_t)sp->n * DF_WSTRIDE;
    if (sp->dense && p == DF_P && smem <= 200 * 1024 && reinterpret_cast<uintptr_t>(c) % 8 == 0) {
        static bool attr_set[64] = {false};
        int dev = 0;
        GCNK_CUDA(cudaGetDevice(&dev));
        if (!attr_set[dev]) {
            GCNK_CUDA(cudaFuncSetAttribute(dense_fw16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[dev] = true;
        }
        // persistent: CTAs sized so that at least two fit per SM at Reddit's n=602 (48 KB of W each)
        const int per_sm = std::max(1, std::min(4, (int)((220 * 1024) / std::max<size_t>(smem, 1))));
        const int n_quads = (sp->m + DF_ROWS - 1) / DF_ROWS;
        const int grid = std::max(1, std::min(sm_count() * per_sm, (n_quads + 7) / 8));
        dense_fw16_kernel<<<grid, 256, smem, st>>>(values, b, c, sp->m, sp->n, drop_bits, drop_scale, row_scale);
        GCNK_LAUNCHED();
        return GCNK_OK;
    }
    if (p <= 1) return launch_fw_generic<1>(sp, values, b, c, p, drop_bits, drop_scale, row_scale, st);
    if (p <= 2) return launch_fw_generic<2>(sp, values, b, c, p, drop_bits, drop_scale, row_scale, st);
    if (p <= 4) return launch_fw_generic<4>(sp, values, b, c, p, drop_bits, drop_scale, row_scale, st);
    if (p <= 8) return launch_fw_generic<8>(sp, values, b, c, p, drop_bits, drop_scale, row_scale, st);
    if (p <= 16) return launch_fw_generic<16>(sp, values, b, c, p, drop_bits, drop_scale, row_scale, st);
    return launch_fw_generic<32>(sp, values, b, c, p, drop_bits, drop_scale, row_scale, st);
}

int gcnk_dense_transform(const float *x, int m, int n, const float *w, float *c, int p, const uint32_t *drop_bits,
                         float drop_scale, const float *row_scale, int relu, gcnk_stream_t stream) {
    GCNK_REQUIRE(x && w && c && m >= 0 && n > 0 && p > 0, "bad arguments");
    if (m == 0) return GCNK_OK;
    if (!dense_tc_supported(n, p) || reinterpret_cast<uintptr_t>(c) % 8 || reinterpret_cast<uintptr_t>(x) % 8) {
        set_error("gcnk_dense_transform: needs p == 16, an even n >= 8 and 8-byte aligned buffers (got n=%d p=%d)", n, p);
        return GCNK_EUNSUPPORTED;
    }
    return dense_fw16_tc(x, w, c, m, n, drop_bits, (int64_t)m * n, drop_scale, row_scale, relu, S(stream));
}

int gcnk_spmm_bw(gcnk_spmat *sp, const float *values, const float *c_grad, float *b_grad, int p, const uint32_t *drop_bits,
                 float drop_scale, gcnk_stream_t stream) {
    GCNK_REQUIRE(sp && values && c_grad && b_grad && p > 0, "bad arguments");
    cudaStream_t st = S(stream);
    const int n = sp->n;
    if (sp->dense && dense_tc_supported(n, p) && sp->m > 0 && reinterpret_cast<uintptr_t>(values) % 8 == 0) {
        const size_t need = dense_bw16_tc_parts(sp->m) * (size_t)n * p;
        if (sp->partial_elems < need) {
            GCNK_CUDA(cudaStreamSynchronize(st));
            if (sp->partials) GCNK_CUDA(cudaFree(sp->partials));
            sp->partials = nullptr; sp->partial_elems = 0;
            GCNK_CUDA(cudaMalloc(&sp->partials, sizeof(float) * need));
            sp->partial_elems = need;
        }
        return dense_bw16_tc(values, c_grad, b_grad, sp->partials, sp->m, n, drop_bits, sp->nnz, drop_scale, st);
    }
    if (sp->dense && p == DB_P && n <= 256 * DB_FPT && sp->m > 0) {
        const int ctas = std::max(1, std::min(sm_count() * 2, (sp->m + 63) / 64));
        const int rows_per_cta = (sp->m + ctas - 1) / ctas;
        const int parts = (sp->m + rows_per_cta - 1) / rows_per_cta;
        const size_t need = (size_t)parts * n * DB_P;
        if (sp->partial_elems < need) {
            GCNK_CUDA(cudaStreamSynchronize(st));
            if (sp->partials) GCNK_CUDA(cudaFree(sp->partials));
            sp->partials = nullptr; sp->partial_elems = 0;
            GCNK_CUDA(cudaMalloc(&sp->partials, sizeof(float) * need));
            sp->partial_elems = need;
        }
        dense_bw16_kernel<<<parts, 256, 0, st>>>(values, c_grad, sp->partials, sp->m, n, rows_per_cta, drop_bits, drop_scale);
        GCNK_LAUNCHED();
        const int elems = n * DB_P;
        reduce_partials_kernel<<<(elems + 255) / 256, 256, 0, st>>>(sp->partials, b_grad, elems, parts);
        GCNK_LAUNCHED();
        return GCNK_OK;
    }
    const int LPsel = p <= 1 ? 1 : p <= 2 ? 2 : p <= 4 ? 4 : p <= 8 ? 8 : p <= 16 ? 16 : 32;
    const int cols_per_block = 256 / LPsel;
    const int grid = std::max(1, std::min((n + cols_per_block - 1) / cols_per_block, sm_count() * 32));
    if (sp->dense) {
#define GCNK_BW_DENSE(LP) spmm_bw_dense_generic_kernel<LP, 4><<<grid, 256, 0, st>>>(values, c_grad, b_grad, sp->m, n, p, drop_bits, drop_scale)
        switch (LPsel) {
        case 1: GCNK_BW_DENSE(1); break; case 2: GCNK_BW_DENSE(2); break; case 4: GCNK_BW_DENSE(4); break;
        case 8: GCNK_BW_DENSE(8); break; case 16: GCNK_BW_DENSE(16); break; default: GCNK_BW_DENSE(32); break;
        }
#undef GCNK_BW_DENSE
        GCNK_LAUNCHED();
        return GCNK_OK;
    }
    if (!sp->csc_ptr) {
        const int rc = build_csc(sp, st);
        if (rc) return rc;
    }
#define GCNK_BW_CSC(LP) spmm_bw_kernel<LP, 4><<<grid, 256, 0, st>>>(sp->csc_ptr, sp->csc_row, sp->csc_pos, values, c_grad, b_grad, n, p, drop_bits, drop_scale)
    switch (LPsel) {
    case 1: GCNK_BW_CSC(1); break; case 2: GCNK_BW_CSC(2); break; case 4: GCNK_BW_CSC(4); break;
    case 8: GCNK_BW_CSC(8); break; case 16: GCNK_BW_CSC(16); break; default: GCNK_BW_CSC(32); break;
    }
#undef GCNK_BW_CSC
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // extern "C"
