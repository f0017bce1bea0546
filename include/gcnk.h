/* include/gcnk.h — the C ABI of libgcnk.so: hand-written sm_100a kernels for the full-batch GCN
 * training path of hengdashi/cuda_gcn, one entry point per kernel row of the reference
 * (src/cuda/cuda_kernel.cuh:20-90) plus the fused variants the B200 plan uses.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers unless a parameter is named h_* (host), sizes as int / int64_t,
 *     a stream as `gcnk_stream_t` (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - every function returns 0 on success, a positive cudaError_t on a CUDA failure, or a negative
 *     GCNK_E* code on a bad argument.  gcnk_last_error() returns a message for the calling thread.
 *     (The reference's convention is CUDA_CHECK -> print + exit, cuda_kernel.cuh:11-18; the C++ host
 *     layer in cuda_gcn_b200/host keeps that convention on top of these return codes.)
 *   - all launches are asynchronous on `stream`; nothing here synchronises unless it says so.
 *   - outputs are WRITTEN, never accumulated into: callers need no zero()/zero_grad() memsets
 *     (the reference needs them because its kernels `+=` into global memory, cuda_module.cu:11,27,45).
 *   - floating point is fp32 (as the reference), indices int32, mask storage is bit-packed uint32.
 *   - there is no CPU fallback anywhere behind this interface.
 */
#ifndef GCNK_H
#define GCNK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *gcnk_stream_t;

enum { GCNK_OK = 0, GCNK_EINVAL = -1, GCNK_ENOMEM = -2, GCNK_EUNSUPPORTED = -3, GCNK_ENODEVICE = -4, GCNK_EASYNC = -5 };

/* ---- library / device ------------------------------------------------------------------------ */
int         gcnk_version(void);                        /* 10000*major + 100*minor + patch */
const char *gcnk_last_error(void);
int         gcnk_device_count(int *count);             /* GCNK_ENODEVICE (and *count = 0) without a GPU */
int         gcnk_set_device(int device);
int         gcnk_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem, int *l2_bytes);
/* The TMA / tcgen05 pipeline kernels bound every mbarrier wait (~1 s) and raise a per-device flag instead of hanging the
 * GPU.  gcnk_async_error synchronises `stream` and returns GCNK_EASYNC (clearing the flag) if any of them timed out since
 * the last call; gcnk_async_error_flag exposes the device int so that a caller can read it with its own result copies. */
int         gcnk_async_error(gcnk_stream_t stream);
int         gcnk_async_error_flag(const int **d_flag);
int64_t     gcnk_launch_count(void);                   /* kernels launched by this library so far (process-wide) */

/* ---- memory, streams, events (thin, so a C/C++/ctypes host needs no other CUDA binding) -------- */
int gcnk_malloc(void **ptr, size_t bytes);
int gcnk_free(void *ptr);
int gcnk_malloc_host(void **ptr, size_t bytes);        /* pinned */
int gcnk_free_host(void *ptr);
int gcnk_memcpy_h2d(void *dst, const void *h_src, size_t bytes, gcnk_stream_t stream);
int gcnk_memcpy_d2h(void *h_dst, const void *src, size_t bytes, gcnk_stream_t stream);
int gcnk_memcpy_d2d(void *dst, const void *src, size_t bytes, gcnk_stream_t stream);
int gcnk_memset(void *dst, int value, size_t bytes, gcnk_stream_t stream);
int gcnk_stream_create(gcnk_stream_t *stream);
int gcnk_stream_create_low_priority(gcnk_stream_t *stream);   /* its CTAs are scheduled after those of other streams */
int gcnk_stream_destroy(gcnk_stream_t stream);
int gcnk_stream_sync(gcnk_stream_t stream);
int gcnk_device_sync(void);
int gcnk_event_create(void **event);
int gcnk_event_destroy(void *event);
int gcnk_event_record(void *event, gcnk_stream_t stream);
int gcnk_event_sync(void *event);
int gcnk_event_elapsed_ms(void *start, void *stop, float *ms);
int gcnk_stream_wait_event(gcnk_stream_t stream, void *event);   /* work queued on stream after this waits for event */
int gcnk_flush_l2(gcnk_stream_t stream);               /* writes a >L2 scratch buffer (bench hygiene) */

/* ---- graph preparation ------------------------------------------------------------------------
 * Replaces the per-edge work the reference redoes in every launch (cuda_kernel.cu:133-136: two
 * indptr loads, an int product, sqrtf and a division per edge per thread): the degree vector
 * d^-1/2 and a static, degree-balanced row->warp schedule are computed once per graph.
 * `d_indptr`/`d_indices` are borrowed and must outlive the handle.  n_cols >= n is the number of
 * rows the gather source has (n_cols > n for a row partition whose column ids are global);
 * d_dinv_global (may be NULL when n_cols == n) supplies d^-1/2 for all n_cols columns. */
typedef struct gcnk_graph gcnk_graph;
int gcnk_graph_create(gcnk_graph **g, const int *d_indptr, const int *d_indices, int n, int64_t nnz,
                      int n_cols, const float *d_dinv_global, gcnk_stream_t stream);
/* A view of `base` (which must outlive it) for passes that need only part of the product:
 *   d_row_keep (int[n] flags, NULL = all rows): only rows with a non-zero flag are scheduled; the other rows of
 *               the output are left untouched (e.g. logits are only needed for the labelled rows of the split);
 *   d_col_keep (int[n_cols] flags, NULL = all): entries whose column flag is 0 are dropped (e.g. the rows of the
 *               gathered matrix that are known to be zero: the loss gradient of unlabelled nodes).
 * Degrees — hence d^-1/2 of rows and columns — remain those of the base graph.  gcnk_graph_stats on a view
 * reports the entries its launches touch.  `base` may itself be a column-filtered view when only d_row_keep is given (the
 * new view shares that view's CSR, which must outlive it).  The filtering runs on the device. */
int gcnk_graph_create_view(gcnk_graph **view, const gcnk_graph *base, const int *d_row_keep, const int *d_col_keep,
                           gcnk_stream_t stream);
int gcnk_graph_destroy(gcnk_graph *g);
int gcnk_graph_dinv(const gcnk_graph *g, const float **d_dinv);   /* [n] d^-1/2 of the local rows */
int gcnk_graph_stats(const gcnk_graph *g, int *n, int64_t *nnz, int *max_degree, int *is_symmetric, int *n_bins);

/* ---- GraphSum: K6/K7, cuda_kernel.cu:126-162 (CPU: module.cpp:83-119) ---------------------------
 * out[s,:] = sum_{d in row s} in[d,:] / sqrt(deg(s)*deg(d)).  Forward and backward are the same
 * operation on (data) resp. (grad) buffers, exactly as in the reference.  `in` has n_cols rows.
 * Widths above 4 that are not a multiple of 4 (41, 47: the class widths) run through rows padded to a 16-byte pitch inside the
 * handle's scratch buffer, so that the gather moves them with 128-bit loads. */
int gcnk_graphsum(const gcnk_graph *g, const float *in, float *out, int dim, gcnk_stream_t stream);
/* gcnk_graphsum keeps a [n_cols x dim] pre-scaled copy of its input inside the handle for the next call; this frees it
 * (synchronise first): worth it after a one-off wide call such as A_hat*X at dim 602 (561 MB at Reddit shape). */
int gcnk_graph_release_scratch(gcnk_graph *g);

/* Building blocks of the fused plan (all row-major [rows x dim], `scaled` = already multiplied by
 * d^-1/2 of its own row so the gather needs no per-edge coefficient):
 *   gcnk_scale_rows:       out[r,:] = dinv[r] * in[r,:]
 *   gcnk_gather_plain:     out[s,:] = dinv[s] * sum_d in_scaled[d,:]
 *   gcnk_gather_relu_drop: a = dinv[s]*sum_d in_scaled[d,:]; h = relu(a); keep = drop bit (NULL: keep all);
 *                          out_scaled[s,:] = dinv[s] * (keep ? h*scale : 0); mask bit = (a>0) & keep
 *                          [fuses K6 + K10 + K12 + the pre-scale for the next gather]
 *   gcnk_gather_mask:      a = dinv[s]*sum_d in_scaled[d,:]; out_scaled[s,:] = dinv[s] * (mask bit ? a*scale : 0)
 *                          [fuses K7 + K13 + K11 + the pre-scale for the next gather]
 * drop_bits is the flat keep stream of gcnk_dropout_mask: bit i*dim+j belongs to element (i,j).
 * mask_bits (written by gather_relu_drop, read by gather_mask) is row-padded so that concurrent rows
 * never share a word: row i starts at bit i*gcnk_mask_row_stride_bits(dim) (8, 16, or dim rounded up
 * to 32), i.e. it holds n*stride/32 words (rounded up). */
int gcnk_mask_row_stride_bits(int dim);
/* Tuning knob of the dim 12 and dim 16 gather (how a warp fetches its edge indices): 0 = one coalesced index load + warp
 * shuffles, 1 = aligned int4 index reads (32 registers, 8 CTAs/SM), 2 = int4 index reads with a chunk's four row reads in
 * flight together (40 registers, 6 CTAs/SM), 3 = two chunks = eight row reads in flight (64 registers, 4 CTAs/SM).
 * 1..3 need a 16-byte-aligned indices array allocated with its length rounded up to 4 entries (checked at
 * gcnk_graph_create; otherwise 0 is used) and add a row's entries in a different (still fixed) order.  v in 0..3
 * selects, anything else only queries; returns the previous value.  Also settable with GCNK_GATHER_IDX4 in the
 * environment; set it before the graph is created, the static row schedule is sized for the variant's occupancy. */
int gcnk_gather_variant(int v);
int gcnk_scale_rows(const float *d_dinv, const float *in, float *out, int rows, int dim, gcnk_stream_t stream);
int gcnk_gather_plain(const gcnk_graph *g, const float *in_scaled, float *out, int dim, gcnk_stream_t stream);
int gcnk_gather_relu_drop(const gcnk_graph *g, const float *in_scaled, float *out_scaled, const uint32_t *drop_bits,
                          uint32_t *mask_bits, float scale, int dim, gcnk_stream_t stream);
int gcnk_gather_mask(const gcnk_graph *g, const float *in_scaled, float *out_scaled, const uint32_t *mask_bits,
                     float scale, int dim, gcnk_stream_t stream);

/* ---- SparseMatmul: K4/K5, cuda_kernel.cu:100-122 (CPU: module.cpp:47-77) -------------------------
 * gcnk_spmat wraps the CSR index of the feature matrix X[m x n] (borrowed device arrays) and adds,
 * once, what the kernels need: a transposed (CSC) view so the backward is a deterministic gather
 * (the reference's backward kernel races on b_grad), and detection of the "dense
 * features" layout (every row stores exactly n entries with column ids 0..n-1, e.g. Reddit as
 * written by reddit_preprocess.py:161-167), for which the index array is never read.
 * fw: c[m x p] = X * b[n x p];   bw: b_grad[n x p] = X^T * c_grad.   values[nnz] are X's entries.
 * Optional fusions (pass NULL / 1.0f to disable):
 *   drop_bits: bit jj set = keep value jj; kept values are multiplied by drop_scale on read (K12
 *              fused; the pristine values are never modified, so no per-pass H2D restore is needed);
 *   row_scale: c[i,:] *= row_scale[i] on write (the d^-1/2 pre-scale for the following gather). */
typedef struct gcnk_spmat gcnk_spmat;
int gcnk_spmat_create(gcnk_spmat **sp, const int *d_indptr, const int *d_indices, int m, int n, int64_t nnz,
                      gcnk_stream_t stream);
int gcnk_spmat_destroy(gcnk_spmat *sp);
int gcnk_spmat_is_dense(const gcnk_spmat *sp, int *is_dense);
int gcnk_spmm_fw(const gcnk_spmat *sp, const float *values, const float *b, float *c, int p,
                 const uint32_t *drop_bits, float drop_scale, const float *row_scale, gcnk_stream_t stream);
int gcnk_spmm_bw(gcnk_spmat *sp, const float *values, const float *c_grad, float *b_grad, int p,
                 const uint32_t *drop_bits, float drop_scale, gcnk_stream_t stream);
/* The same forward for a plain row-major dense matrix x[m x n] (no index at all), with an optional ReLU in the
 * epilogue: c = row_scale (.) relu?(drop(x) * w).  Used with x = A_hat*X precomputed once, which turns the
 * dropout-free layer-1 forward A_hat*(X*W1) into one streaming pass (A_hat*X)*W1 with no gather.
 * Tensor-core path only (p == 16, n even): other shapes return GCNK_EUNSUPPORTED. */
int gcnk_dense_transform(const float *x, int m, int n, const float *w, float *c, int p, const uint32_t *drop_bits,
                         float drop_scale, const float *row_scale, int relu, gcnk_stream_t stream);

/* TMA-staged variants of the same transform and of its weight gradient (w_grad[n x p] = drop(x)^T * g[m x p]) for a
 * PACKED matrix: row pitch `ld` floats with ld*4 a multiple of 16 bytes (the reference's 602-float rows are not TMA-
 * addressable).  gcnk_dense_pack makes the packed copy (zero padded); the keep bits stay in the reference's flat order
 * (bit row*n + col of the unpadded matrix).  p == 16 only; GCNK_EUNSUPPORTED otherwise.  The backward needs
 * gcnk_dense_transform_bw_workspace(m, n) bytes of scratch (per-CTA partial sums, reduced in a fixed order). */
int gcnk_dense_pack(const float *x, int m, int n, float *xp, int ld, gcnk_stream_t stream);
int gcnk_dense_transform_ld(const float *xp, int ld, int m, int n, const float *w, float *c, int p, const uint32_t *drop_bits,
                            float drop_scale, const float *row_scale, int relu, gcnk_stream_t stream);
size_t gcnk_dense_transform_bw_workspace(int m, int n);
int gcnk_dense_transform_bw_ld(const float *xp, int ld, int m, int n, const float *g, float *w_grad, int p, const uint32_t *drop_bits,
                               float drop_scale, float *workspace, size_t workspace_bytes, gcnk_stream_t stream);

/* The same transform and weight gradient on the 5th-generation tensor cores (csrc/matmul_tc.cu: tcgen05.mma kind::tf32 with
 * 3xTF32 split products, accumulators in TMEM, X tiles staged by TMA; Dropout applied while a tile is split in shared
 * memory).  xp: packed as for gcnk_dense_transform_ld; any p that is a multiple of 4 up to 256.  GCNK_EUNSUPPORTED when
 * the shape does not qualify (fewer than ~2,000 rows, GCNK_NO_TCGEN05=1): callers fall back to the _ld forms. */
int gcnk_dense_transform_tc(const float *xp, int ld, int m, int n, const float *w, float *c, int p, const uint32_t *drop_bits,
                            float drop_scale, const float *row_scale, int relu, gcnk_stream_t stream);
size_t gcnk_dense_transform_bw_tc_workspace(int m, int n, int p);
int gcnk_dense_transform_bw_tc(const float *xp, int ld, int m, int n, const float *g, float *w_grad, int p, const uint32_t *drop_bits,
                               float drop_scale, float *workspace, size_t workspace_bytes, gcnk_stream_t stream);

/* ---- Matmul: K1/K2/K3, cuda_kernel.cu:6-96 (CPU: module.cpp:11-42) -------------------------------- */
int gcnk_matmul_fw(const float *a, const float *b, float *c, int m, int n, int p, gcnk_stream_t stream);   /* c = a*b       */
int gcnk_matmul_bw_a(const float *c_grad, const float *b, float *a_grad, int m, int n, int p, gcnk_stream_t stream); /* a_grad = c_grad * b^T */
int gcnk_matmul_bw_b(const float *a, const float *c_grad, float *b_grad, int m, int n, int p,
                     float *workspace, size_t workspace_bytes, gcnk_stream_t stream);                      /* b_grad = a^T * c_grad */
size_t gcnk_matmul_bw_b_workspace(int m, int n, int p);

/* The same three products for operands inside padded buffers (explicit row pitches lda/ldb/ldc, in floats):
 *   nn: C[m x n] = A[m x k] * B[k x n], optional row_scale (C[i,:] *= row_scale[i]: the GraphSum pre-scale)   Matmul::forward
 *   nt: C[m x n] = A[m x k] * B[n x k]^T                                                                      dA = dC * B^T
 *   tn: C[ka x n] = A[m x ka]^T * B[m x n], contraction over the node dimension m (split across CTAs, partial tiles summed
 *       in a fixed order); needs gcnk_matmul_tn_workspace(m, ka, n) bytes                                     dB = A^T * dC
 * Large shapes with 16-byte-aligned pitches run on the tensor cores (tcgen05.mma kind::tf32 with 3xTF32 split products,
 * accumulators in TMEM, operands staged by TMA: csrc/matmul_tc.cu); anything else on the fp32 SIMT kernel.  GCNK_NO_TCGEN05=1
 * forces the latter. */
int gcnk_matmul_nn(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int k, int n,
                   const float *row_scale, gcnk_stream_t stream);
int gcnk_matmul_nt(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int k, int n, gcnk_stream_t stream);
size_t gcnk_matmul_tn_workspace(int m, int ka, int n);
int gcnk_matmul_tn(const float *a, int lda, const float *b, int ldb, float *c, int ldc, int m, int ka, int n,
                   float *workspace, size_t workspace_bytes, gcnk_stream_t stream);

/* ---- ReLU: K10/K11, cuda_kernel.cu:204-219 (CPU: module.cpp:175-194) ------------------------------- */
int gcnk_relu_fw(float *x, uint32_t *mask_bits, int64_t n, int training, gcnk_stream_t stream);  /* mask untouched when !training */
int gcnk_relu_bw(float *grad, const uint32_t *mask_bits, int64_t n, gcnk_stream_t stream);

/* ---- Dropout: K12/K13, cuda_kernel.cu:223-240 (CPU: module.cpp:207-233) -----------------------------
 * The reference CPU engine draws one xorshift128+ value per element from a single global stream
 * (rand.cpp:17-28) and keeps element i iff (int)draw >= int(p*0x7fffffff).  gcnk_rng reproduces THAT
 * stream bit-for-bit in parallel: xorshift128+ is linear over GF(2), so every thread jumps to its own
 * offset with precomputed powers of the transition matrix.  (This replaces the reference GPU path's
 * racy shared cuRAND states, cuda_kernel.cu:229.)  State lives on the host, like rand_state[2]. */
typedef struct gcnk_rng gcnk_rng;
int gcnk_rng_create(gcnk_rng **rng, uint64_t s0, uint64_t s1);
int gcnk_rng_destroy(gcnk_rng *rng);
int gcnk_rng_seed(gcnk_rng *rng, long seed);                 /* srand(seed); rand(); rand() exactly as rand.cpp:6-15 */
int gcnk_rng_get_state(const gcnk_rng *rng, uint64_t *h_state2);
int gcnk_rng_set_state(gcnk_rng *rng, uint64_t s0, uint64_t s1);
int gcnk_rng_skip(gcnk_rng *rng, uint64_t n_draws);          /* advance the host state by n draws, O(log n) */
int gcnk_rng_next_host(gcnk_rng *rng, uint32_t *h_out, int64_t n);   /* host-side draws (Glorot init), advances */
/* keep bits for the next n draws (bit i set = keep), advances the stream by n.  p is the dropout rate.  From 2^20 draws on the
 * bit-sliced kernels run (csrc/rng_bitsliced.cuh: 32 streams per thread); the bits are the reference's either way. */
int gcnk_dropout_mask(gcnk_rng *rng, uint32_t *keep_bits, int64_t n, float p, gcnk_stream_t stream);
int gcnk_dropout_apply(float *x, const uint32_t *keep_bits, int64_t n, float p, gcnk_stream_t stream);  /* fw and bw: x *= keep ? 1/(1-p) : 0 */

/* ---- CrossEntropyLoss: K8/K9 + thrust reductions, cuda_kernel.cu:166-200, cuda_module.cu:121-146
 * (CPU: module.cpp:124-161) and GCN::get_accuracy (gcn.cpp:83-96) in the same pass.
 * logits[n x c] are shifted in place by the row max for labelled rows (as the reference does);
 * grad (NULL when !training) receives (softmax - onehot)/count, zeros for unlabelled rows.
 * d_result[4] (device floats/ints, written): {loss (mean over labelled rows), count, wrong, 0}. */
typedef struct { float loss; int count; int wrong; int pad; } gcnk_ce_result;
int gcnk_softmax_ce(float *logits, const int *truth, float *grad, int n, int c, int training,
                    gcnk_ce_result *d_result, float *workspace, size_t workspace_bytes, gcnk_stream_t stream);
size_t gcnk_softmax_ce_workspace(int n, int c);
int gcnk_accuracy(const float *logits, const int *truth, int n, int c, int *d_wrong_total2, gcnk_stream_t stream);
int gcnk_set_truth(int *truth, const int *split, const int *label, int current_split, int n, gcnk_stream_t stream);  /* K16 */

/* ---- Adam: K15, cuda_kernel.cu:270-281 (CPU: optim.cpp:24-37) + L2 penalty (gcn.cpp:98-105) ---------
 * One launch for all tensors.  step_size = lr*sqrtf(1-beta2^t)/(1-beta1^t) is computed by the caller
 * in fp32 as the reference does.  m,v must start zeroed (the reference GPU path forgets).
 * d_sumsq (may be NULL): receives sum(w^2) of tensor 0 AFTER the update (the next pass's L2 penalty). */
typedef struct { float *data; const float *grad; float *m; float *v; int size; int decay; } gcnk_adam_tensor;
int gcnk_adam_step(const gcnk_adam_tensor *h_tensors, int count, float step_size, float beta1, float beta2,
                   float eps, float weight_decay, float *d_sumsq, gcnk_stream_t stream);
int gcnk_sum_squares(const float *w, int64_t n, float *d_out, gcnk_stream_t stream);

/* ---- fused layer 2 (row-local): Matmul (K1-K3) + CE (K8/K9) + accuracy in one pass ------------------
 * For each row s with P[s,:] = (A_hat * H1)[s,:] already aggregated (hidden dim h):
 *   logits = P[s,:] * W2  (h x c);  labelled rows (split[s]==current_split): loss, wrong;
 *   training: dlogits = (softmax-onehot)/count;  G_scaled[s,:] = dinv[s] * (dlogits * W2^T);
 *             W2_grad = P^T * dlogits  (= H1^T * A_hat * dlogits for a symmetric A_hat)
 * This is the algebraic re-ordering A_hat*(H1*W2) -> (A_hat*H1)*W2 (hard part 3): the
 * gather runs at width h instead of c.  logits_out may be NULL.  count = #labelled rows (static).
 * d_result as gcnk_softmax_ce.  Requires h*c <= 4096.  On return the first four floats of `workspace` hold the
 * raw sums {sum of loss terms, count, wrong, 0} (counts as floats, exact below 2^24) for cross-rank reduction. */
int gcnk_layer2_fused(const float *P, const float *W2, const int *split, const int *label, int current_split,
                      int n, int h, int c, int training, int count, const float *d_dinv,
                      float *G_scaled, float *W2_grad, float *logits_out, gcnk_ce_result *d_result,
                      float *workspace, size_t workspace_bytes, gcnk_stream_t stream);
size_t gcnk_layer2_workspace(int n, int h, int c);
/* The same, also storing every labelled row's loss term: loss_terms[term_index ? term_index[s] : s] = log(sum exp) - logit
 * [truth] (without term_index the rows without a label get 0).  For gcnk_sequential_sum. */
int gcnk_layer2_fused_terms(const float *P, const float *W2, const int *split, const int *label, int current_split,
                            int n, int h, int c, int training, int count, const float *d_dinv,
                            float *G_scaled, float *W2_grad, float *logits_out, gcnk_ce_result *d_result,
                            float *workspace, size_t workspace_bytes, float *loss_terms, const int *term_index,
                            gcnk_stream_t stream);
/* d_out[0] = ((((0 + terms[0]) + terms[1]) + ...) + terms[n-1]) [/ divide_by if non-zero] in fp32, bit for bit what a scalar
 * loop computes — the reference accumulates its loss that way (module.cpp:125-143), and at 10^5 labelled rows the rounding
 * of that loop is larger than the parity tolerance — but evaluated block-parallel: one CTA rounds 4,096 terms at a time in the
 * running sum's binade, one warp commits the blocks in order (exact integer additions; anything unusual is replayed term by term).
 * Optional d_wait_flags as in gcnk_gather_wait_next (row-partitioned runs: the terms of the other ranks arrive by push). */
int gcnk_sequential_sum(const float *terms, int n, float *d_out, float divide_by, const int *d_wait_flags, int n_flags, int skip,
                        int wait_value, int *d_err, gcnk_stream_t stream);

/* ---- row-local pieces of the wide-hidden plan (hidden*classes too large for gcnk_layer2_fused; csrc/wide.cu) ------
 * The wide plan computes layer 1 as (A_hat * drop(X)) * W1 (gather at the input width, result reused for dW1) and layer 2
 * in the reference's order at the class width, stored with pitch ld = classes rounded up to 4 (zero padding columns).
 *   gcnk_drop_scale_rows   out[i,:] = dinv[i] * (keep bit(i*f+j) ? x[i,j]*scale : 0): Dropout forward (module.cpp:207-224) +
 *                          the GraphSum pre-scale, from the pristine features; keep_bits NULL keeps everything (f % 4 == 0)
 *   gcnk_relu_dropout_fw   in place h = (z > 0 && keep) ? z*scale : 0, mask bit = (z > 0) && keep  (K10 + K12)
 *   gcnk_mask_scale_bw     in place g = mask ? g*scale : 0                                         (K13 + K11)
 *   gcnk_pad_cols / gcnk_unpad_cols   [rows x c] <-> [rows x ld] with zero padding columns
 *   gcnk_ce_rows           softmax-CE + accuracy over rows of pitch ld, labelled rows of `current_split` only; training:
 *                          grad_scaled[s,:] = dinv[s] * (softmax - onehot)/count (pitch ld; zero rows for the others);
 *                          workspace as gcnk_layer2_fused (first four floats = raw sums); loss_terms/term_index as
 *                          gcnk_layer2_fused_terms. */
int gcnk_drop_scale_rows(const float *x, int rows, int f, const uint32_t *keep_bits, float scale, const float *d_dinv, float *out,
                         gcnk_stream_t stream);
int gcnk_relu_dropout_fw(float *z, int64_t n, const uint32_t *keep_bits, float scale, uint32_t *mask_bits, gcnk_stream_t stream);
int gcnk_mask_scale_bw(float *g, int64_t n, const uint32_t *mask_bits, float scale, gcnk_stream_t stream);
int gcnk_pad_cols(const float *src, float *dst, int rows, int c, int ld, gcnk_stream_t stream);
int gcnk_unpad_cols(const float *src, float *dst, int rows, int c, int ld, gcnk_stream_t stream);
size_t gcnk_ce_rows_workspace(int n);
int gcnk_ce_rows(const float *logits, int ld, const int *split, const int *label, int current_split, int n, int c, int training,
                 int count, const float *d_dinv, float *grad_scaled, gcnk_ce_result *d_result, float *workspace,
                 size_t workspace_bytes, float *loss_terms, const int *term_index, gcnk_stream_t stream);

/* ---- exchange steps of the row-partitioned engine (NCCL over NVLink; the reference is single-GPU) ------
 * One process (or thread) per GPU.  Rank 0 calls gcnk_comm_unique_id and shares the 128 bytes with the
 * other ranks by any means; every rank then calls gcnk_comm_create.  Errors: 1000 + ncclResult_t. */
typedef struct gcnk_comm gcnk_comm;
int gcnk_comm_unique_id(void *h_id128);
int gcnk_comm_create(gcnk_comm **comm, const void *h_id128, int rank, int world, int device);
int gcnk_comm_destroy(gcnk_comm *comm);
int gcnk_comm_rank(const gcnk_comm *comm, int *rank, int *world);
/* In-place all-gather of row blocks: d_all is [h_row_begin[world] x dim]; on entry rank r's rows
 * [h_row_begin[r], h_row_begin[r+1]) are valid on rank r, on return all rows are valid everywhere. */
int gcnk_comm_allgather_rows(gcnk_comm *comm, float *d_all, const int *h_row_begin, int dim, gcnk_stream_t stream);
/* In-place all-reduce (sum, or max when op_max != 0) of n_bufs device buffers in one group. */
int gcnk_comm_allreduce(gcnk_comm *comm, float *const *d_bufs, const size_t *h_counts, int n_bufs, int op_max,
                        gcnk_stream_t stream);

/* ---- fused all-gather over NVLink peer memory ----------------------------------------------------------
 * With one process per GPU, each rank exports the buffer that holds its gather sources (gcnk_ipc_export, a
 * 64-byte handle), exchanges the handles (gcnk_comm_allgather_bytes) and maps the peers' buffers
 * (gcnk_ipc_import).  gcnk_mirror_next(out, peer_out, n) then makes the NEXT producer launched on `out`
 * (gcnk_spmm_fw / gcnk_dense_transform at p == 16, gcnk_gather_*, gcnk_layer2_fused at h == 16) store every row
 * it writes also at the same offset of each peer_out[i]: the all-gather rides in the producer's epilogue.
 * gcnk_peer_push is the unfused form (a copy kernel) for producers without a mirrored epilogue.
 * gcnk_peer_barrier, launched after the producer, returns (in stream order) once every rank has finished the
 * same step: flag_arrays[r] is rank r's int[world] flag array (own + imported), `value` must increase with
 * every barrier, *d_err is set to 1 if a peer does not arrive within ~2 s. */
int gcnk_ipc_export(const void *dptr, void *h_handle64);
int gcnk_ipc_import(void **dptr, const void *h_handle64);
int gcnk_ipc_release(void *dptr);
int gcnk_comm_allgather_bytes(gcnk_comm *comm, const void *h_send, void *h_recv, int bytes_per_rank);
int gcnk_mirror_next(const float *local_out, float *const *peer_out, int n_peers);
int gcnk_mirror_pending(const float *local_out);   /* 1 (and cancels it) if the registration was not consumed by a producer */
int gcnk_peer_push(const float *local_rows, float *const *peer_rows, int n_peers, size_t n_floats, gcnk_stream_t stream);
int gcnk_peer_barrier(int *const *flag_arrays, int rank, int world, int value, int *d_err, gcnk_stream_t stream);
/* push + barrier in one launch (the last CTA to finish its copies runs the flag exchange); d_counter: a zeroed
 * device unsigned reserved for these calls */
int gcnk_peer_push_barrier(const float *local_rows, float *const *peer_rows, int n_peers, size_t n_floats, int *const *flag_arrays,
                           int rank, int world, int value, int *d_err, unsigned *d_counter, gcnk_stream_t stream);
/* push + SIGNAL, and the matching wait inside the consumer (the form the row-partitioned engine uses by default): the
 * rows go to the peers (the contiguous block, or — halo exchange — only the rows listed per peer: d_row_lists[i] =
 * n rows' indices relative to local_rows, h_row_counts[i] of them, `dim` floats each) and the last CTA to finish
 * stores `value` into peer_flag_slots[i] (an int in peer i's memory); nothing waits in this launch.
 * gcnk_gather_wait_next(d_flags, n, skip, value, d_err) makes the NEXT gcnk_gather_* launched by this thread wait, at
 * its start, until d_flags[r] >= value for every r < n except r == skip (this rank) — so a rank that is ahead keeps
 * running its own work, and no barrier kernel sits between producer and consumer.  A peer that does not arrive within
 * GCN_PEER_TIMEOUT_S seconds (default 60) sets *d_err = 1 instead of hanging the device. */
int gcnk_peer_push_signal(const float *local_rows, float *const *peer_rows, int n_peers, size_t n_floats, const int *const *d_row_lists,
                          const int *h_row_counts, int dim, int *const *peer_flag_slots, int value, unsigned *d_counter, gcnk_stream_t stream);
int gcnk_gather_wait_next(const int *d_flags, int n_flags, int skip, int value, int *d_err);
/* Split aggregation (the overlap of the exchange with the local part of a GraphSum): gcnk_gather_raw over a view that keeps
 * only the columns this rank owns writes the UNSCALED partial row sums; gcnk_gather_init_next(partial) makes the next
 * gcnk_gather_* launch — over the view of the remaining columns — start every row sum from partial[s, :] before it applies
 * its usual epilogue.  Own columns first, remote columns second: a fixed order. */
/* The exchange fused INTO the consuming GraphSum (width 12 / 16): gcnk_graph_rotate(g, lo, hi) re-orders every row of a
 * row-partition's slice graph (before any view of it is created) as [columns this rank owns: lo <= c < hi | higher ranks |
 * lower ranks]; gcnk_gather_exchange_next then makes the NEXT gcnk_gather_* launch on g (or a view of it) one kernel that
 * (a) copies own_rows (this rank's finished rows of the gather source, n_floats) into peer_rows[i] and publishes `value`
 * in peer_flag_slots[i], with a few CTAs, and (b) aggregates with the others, starting with the own columns and waiting
 * for d_wait_flags[r] >= value only when it reaches rank r's columns — the transfer overlaps the aggregation inside
 * one launch.  Replaces gcnk_peer_push_signal + gcnk_gather_wait_next for that launch. */
int gcnk_graph_rotate(gcnk_graph *g, int col_lo, int col_hi, gcnk_stream_t stream);
int gcnk_gather_exchange_next(const float *own_rows, size_t n_floats, float *const *peer_rows, int n_peers, int *const *peer_flag_slots,
                              const int *d_wait_flags, int rank, int world, int value, unsigned *d_counter, int *d_err);
int gcnk_gather_raw(const gcnk_graph *g, const float *in_scaled, float *out_raw, int dim, gcnk_stream_t stream);
int gcnk_gather_init_next(const float *d_partial);
/* Deterministic sum all-reduce over peer memory for small vectors (weight gradients, scalars): every rank writes
 * its n_segs segments, packed, into slot[rank] of every rank's exchange area (slot_areas[r] = rank r's area of
 * world * slot_floats floats), passes the barrier, and sums the slots in rank order back into the segments —
 * the same order on every rank, so the replicated weights stay bit-identical. */
int gcnk_peer_allreduce(float *const *d_segs, const size_t *h_counts, int n_segs, float *const *slot_areas, size_t slot_floats,
                        int *const *flag_arrays, int rank, int world, int value, int *d_err, unsigned *d_counter, gcnk_stream_t stream);

/* ---- host-side, bit-exact integer work --------------------------------------------------------------
 * Contiguous nnz-balanced row partition of a CSR: h_row_begin[parts+1] receives the cuts;
 * slice k is rows [h_row_begin[k], h_row_begin[k+1]) with column ids left global. */
int gcnk_partition_rows(const int *h_indptr, int n, int parts, int *h_row_begin);

#ifdef __cplusplus
}
#endif
#endif /* GCNK_H */
