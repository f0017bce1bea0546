// layer2.cu — the row-local half of GCN layer 2, fused: Matmul forward (K1), softmax cross-entropy
// (K8/K9 + thrust reductions), accuracy (CUDAGCN::get_accuracy, a D2H copy + host loop in the
// reference, cuda_gcn.cu:100-120), Matmul backward-A (K2) and backward-B (K3) in ONE pass over the
// aggregated hidden rows.  Reference semantics: src/seq/module.cpp:11-42,124-161, src/seq/gcn.cpp:83-96.
//
// It relies on the re-ordering  A_hat*(H1*W2) = (A_hat*H1)*W2  (hard part 3): the gather that
// precedes this kernel runs at the hidden width h (16) instead of the class count c (41), and the
// logits [n x c] (38 MB at Reddit shape) never have to exist in HBM.  For the backward,
//     dH1 = A_hat * (dlogits * W2^T)          -> this kernel emits G = dinv (.) (dlogits * W2^T), pre-scaled
//     dW2 = H1^T * (A_hat * dlogits) = P^T * dlogits   for a symmetric A_hat, P = A_hat*H1 (already here)
// so the class-width backward gather of the reference disappears as well.
//
// One warp per row (static row -> warp map => fixed summation order): lane l owns classes l, l+32, ...
#include <algorithm>

#include "common.cuh"

using namespace gcnk;

namespace {

constexpr int L2_THREADS = 256, L2_WARPS = L2_THREADS / 32, CPL_MAX = 4;   // c <= 128

struct L2Partial { float loss; int count; int wrong; int pad; };

__global__ void __launch_bounds__(L2_THREADS) layer2_kernel(const float *__restrict__ P, const float *__restrict__ W2,
                                                             const int *__restrict__ split, const int *__restrict__ label,
                                                             int current_split, int n, int h, int c, int training, float count_f,
                                                             const float *__restrict__ dinv, float *__restrict__ G,
                                                             float *__restrict__ logits_out, L2Partial *__restrict__ ce_partials,
                                                             float *__restrict__ dw_partials, float *__restrict__ terms,
                                                             const int *__restrict__ term_index) {
    extern __shared__ __align__(16) float smem[];
    float *sW = smem;                               // [h][c]
    float *sDl = sW + h * c;                        // [warps][c]   dlogits of the warp's current row
    float *sP = sDl + L2_WARPS * c;                 // [warps][h]   P row
    float *sAcc = sP + L2_WARPS * h;                // [warps][h*c] per-warp dW2 accumulators (training only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < h * c; i += L2_THREADS) sW[i] = W2[i];
    if (training)
        for (int i = threadIdx.x; i < L2_WARPS * h * c; i += L2_THREADS) sAcc[i] = 0.f;
    __syncthreads();
    float *dl = sDl + warp * c, *pr = sP + warp * h, *acc = sAcc + warp * h * c;

    float loss = 0.f;
    int count = 0, wrong = 0;
    const int total_warps = gridDim.x * L2_WARPS;
    for (int s = blockIdx.x * L2_WARPS + warp; s < n; s += total_warps) {
        for (int k = lane; k < h; k += 32) pr[k] = P[(size_t)s * h + k];
        __syncwarp();
        float lg[CPL_MAX];
#pragma unroll
        for (int t = 0; t < CPL_MAX; t++) {
            const int cls = lane + 32 * t;
            float v = 0.f;
            if (cls < c)
                for (int k = 0; k < h; k++) v = fmaf(pr[k], sW[k * c + cls], v);
            lg[t] = v;
            if (logits_out && cls < c) logits_out[(size_t)s * c + cls] = v;
        }
        const int truth = split[s] == current_split ? label[s] : -1;        // set_truth (gcn.cpp:78-81)
        if (truth >= 0) {                                                    // warp-uniform
            float mx = -1e30f;
#pragma unroll
            for (int t = 0; t < CPL_MAX; t++) if (lane + 32 * t < c) mx = fmaxf(mx, lg[t]);
            mx = warp_max(mx);
            float ex[CPL_MAX], sum = 0.f;
#pragma unroll
            for (int t = 0; t < CPL_MAX; t++) {
                ex[t] = (lane + 32 * t < c) ? expf(lg[t] - mx) : 0.f;
                sum += ex[t];
            }
            sum = warp_sum(sum);
            float tl = 0.f;                                                  // the truth logit, broadcast from its owner
#pragma unroll
            for (int t = 0; t < CPL_MAX; t++) if (t == truth / 32) tl = lg[t];
            tl = __shfl_sync(FULL, tl, truth % 32);
            bool w = false;
#pragma unroll
            for (int t = 0; t < CPL_MAX; t++) w |= (lane + 32 * t < c) && lg[t] > tl;   // strict (gcn.cpp:88-93)
            w = __any_sync(FULL, w);
            count++;
            wrong += w;
            const float term = logf(sum) - (tl - mx);
            loss += term;
            if (terms && lane == 0) terms[term_index ? term_index[s] : s] = term;
            if (training) {
#pragma unroll
                for (int t = 0; t < CPL_MAX; t++) {
                    const int cls = lane + 32 * t;
                    if (cls < c) {
                        float g = ex[t] / sum;
                        if (cls == truth) g -= 1.0f;
                        g = g / count_f;                                     // grad /= count (module.cpp:156-158)
                        dl[cls] = g;
                        for (int k = 0; k < h; k++) acc[k * c + cls] = fmaf(pr[k], g, acc[k * c + cls]);   // dW2 += P^T dlogits
                    }
                }
                __syncwarp();
                const float di = dinv[s];
                for (int k = lane; k < h; k += 32) {
                    float v = 0.f;
                    for (int cls = 0; cls < c; cls++) v = fmaf(dl[cls], sW[k * c + cls], v);   // dlogits * W2^T
                    G[(size_t)s * h + k] = di * v;
                }
            }
        } else {
            if (training) for (int k = lane; k < h; k += 32) G[(size_t)s * h + k] = 0.f;   // unlabelled rows carry no gradient
            if (terms && !term_index && lane == 0) terms[s] = 0.f;
        }
        __syncwarp();
    }

    // per-CTA partials, combined in warp order
    __shared__ float s_loss[L2_WARPS];
    __shared__ int s_count[L2_WARPS], s_wrong[L2_WARPS];
    if (lane == 0) { s_loss[warp] = loss; s_count[warp] = count; s_wrong[warp] = wrong; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f; int cn = 0, wr = 0;
        for (int w = 0; w < L2_WARPS; w++) { l += s_loss[w]; cn += s_count[w]; wr += s_wrong[w]; }
        ce_partials[blockIdx.x] = L2Partial{l, cn, wr, 0};
    }
    if (training) {
        float *out = dw_partials + (size_t)blockIdx.x * h * c;
        for (int i = threadIdx.x; i < h * c; i += L2_THREADS) {
            float v = 0.f;
            for (int w = 0; w < L2_WARPS; w++) v += sAcc[w * h * c + i];
            out[i] = v;
        }
    }
}


// ---------------------------------------------------------------------------- h == 16 fast path ----
// Same contract as layer2_kernel, specialised for hidden width 16 (the reference default, gcn.cpp:10) and
// c <= 32*CPL classes.  Everything per-row lives in registers: the P row (16 floats, loaded by every lane
// from the same address = one broadcast transaction), the lane's logits, and the lane's slice of the
// W2-gradient accumulator dw[CPL][16] (the generic kernel does 16*c shared-memory read-modify-writes per
// labelled row instead).  Unlabelled rows cost one split[] read (plus zeroing their G row when training).
template <int CPL>
__global__ void __launch_bounds__(L2_THREADS) layer2_h16_kernel(const float *__restrict__ P, const float *__restrict__ W2,
                                                                 const int *__restrict__ split, const int *__restrict__ label,
                                                                 int current_split, int n, int c, int training, float count_f,
                                                                 const float *__restrict__ dinv, float *__restrict__ G,
                                                                 float *__restrict__ logits_out, L2Partial *__restrict__ ce_partials,
                                                                 float *__restrict__ dw_partials, const Mirror mirror,
                                                                 float *__restrict__ terms, const int *__restrict__ term_index) {
    constexpr int H = 16;
    extern __shared__ __align__(16) float smem[];
    float *sAcc = smem;                             // [warps][16*c]  (training only; used once, at the end)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this lane's columns of W2 (classes lane, lane+32) and its slice of the W2 gradient stay in registers
    float w[CPL][H], dw[CPL][H];
#pragma unroll
    for (int t = 0; t < CPL; t++) {
        const int cls = lane + 32 * t;
#pragma unroll
        for (int k = 0; k < H; k++) { w[t][k] = cls < c ? __ldg(W2 + k * c + cls) : 0.f; dw[t][k] = 0.f; }
    }

    float loss = 0.f;
    int count = 0, wrong = 0;
    const int total_warps = gridDim.x * L2_WARPS;
    const float inv_count = 1.0f / count_f;
    // A warp takes blocks of 32 consecutive rows: split[] / label[] are read coalesced (one row per lane), a ballot gives
    // the labelled rows, and only those are processed — an eval pass over a 10 % split touches 10 % of the P rows.
    // The next labelled row's P is in flight while the current one (a ~700-cycle dependent chain) is processed.
    const int n_blocks = (n + 31) / 32;
    for (int blk = blockIdx.x * L2_WARPS + warp; blk < n_blocks; blk += total_warps) {
        const int base = blk * 32, my_row = base + lane;
        const int my_truth = (my_row < n && split[my_row] == current_split) ? label[my_row] : -1;   // set_truth (gcn.cpp:78-81)
        unsigned todo = __ballot_sync(FULL, logits_out ? my_row < n : my_truth >= 0);
        if (terms && !term_index && my_row < n && my_truth < 0) terms[my_row] = 0.f;
        // where this lane's row stores its loss term: read coalesced here, not per row on the dependent chain below
        const int my_slot = (terms && my_truth >= 0) ? (term_index ? term_index[my_row] : my_row) : 0;
        float my_term = 0.f;
        if (training) {
            // rows without a label carry no gradient: zero their G rows, 32 consecutive floats per store
            const unsigned labelled = __ballot_sync(FULL, my_truth >= 0);
#pragma unroll
            for (int j = 0; j < H; j++) {
                const int idx = j * 32 + lane, r = idx >> 4;
                if (base + r < n && !((labelled >> r) & 1u)) {
                    G[(size_t)base * H + idx] = 0.f;
                    mirror_store(mirror, (size_t)base * H + idx, 0.f);
                }
            }
        }
        float4 nq[4] = {};
        if (todo) {
            const float4 *pr = reinterpret_cast<const float4 *>(P + (size_t)(base + __ffs(todo) - 1) * H);
            nq[0] = __ldg(pr); nq[1] = __ldg(pr + 1); nq[2] = __ldg(pr + 2); nq[3] = __ldg(pr + 3);
        }
        while (todo) {
        const int r_in = __ffs(todo) - 1;
        todo &= todo - 1;
        const int s = base + r_in;
        const int truth = __shfl_sync(FULL, my_truth, r_in);                  // warp-uniform
        float p[H];
        p[0] = nq[0].x; p[1] = nq[0].y; p[2] = nq[0].z; p[3] = nq[0].w; p[4] = nq[1].x; p[5] = nq[1].y; p[6] = nq[1].z; p[7] = nq[1].w;
        p[8] = nq[2].x; p[9] = nq[2].y; p[10] = nq[2].z; p[11] = nq[2].w; p[12] = nq[3].x; p[13] = nq[3].y; p[14] = nq[3].z; p[15] = nq[3].w;
        if (todo) {
            const float4 *pr = reinterpret_cast<const float4 *>(P + (size_t)(base + __ffs(todo) - 1) * H);
            nq[0] = __ldg(pr); nq[1] = __ldg(pr + 1); nq[2] = __ldg(pr + 2); nq[3] = __ldg(pr + 3);
        }
        float lg[CPL];
#pragma unroll
        for (int t = 0; t < CPL; t++) {
            const int cls = lane + 32 * t;
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < H; k++) v = fmaf(p[k], w[t][k], v);
            if (logits_out && cls < c) logits_out[(size_t)s * c + cls] = v;
            lg[t] = v;
        }
        if (truth < 0) continue;                                             // logits_out only: the row has no label
        float mx = -1e30f;
#pragma unroll
        for (int t = 0; t < CPL; t++) if (lane + 32 * t < c) mx = fmaxf(mx, lg[t]);
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
        for (int t = 0; t < CPL; t++) wr |= (lane + 32 * t < c) && lg[t] > tl;       // strict (gcn.cpp:88-93)
        wr = __any_sync(FULL, wr);
        count++;
        wrong += wr;
        const float term = logf(sum) - (tl - mx);
        loss += term;
        if (lane == r_in) my_term = term;                                    // stored after the block, one coalesced-ish pass
        if (training) {
            // dlogits of this lane's classes; dW2 += P^T dlogits; partial of dlogits * W2^T over this lane's classes
            float pk[H];
#pragma unroll
            for (int k = 0; k < H; k++) pk[k] = 0.f;
#pragma unroll
            for (int t = 0; t < CPL; t++) {
                const int cls = lane + 32 * t;
                float g = 0.f;
                if (cls < c) {
                    g = ex[t] / sum;
                    if (cls == truth) g -= 1.0f;
                    g = g * inv_count;                                               // grad /= count (module.cpp:156-158)
                }
#pragma unroll
                for (int k = 0; k < H; k++) { dw[t][k] = fmaf(p[k], g, dw[t][k]); pk[k] = fmaf(g, w[t][k], pk[k]); }
            }
            // transposing butterfly: 16 per-lane partials -> lane l holds the warp-wide sum for hidden unit l/2
#pragma unroll
            for (int half = 8, off = 16; off >= 2; half >>= 1, off >>= 1) {
                const bool upper = lane & off;
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const float send = upper ? pk[i] : pk[i + half];
                    const float keep = upper ? pk[i + half] : pk[i];
                    pk[i] = keep + __shfl_xor_sync(FULL, send, off);
                }
            }
            pk[0] += __shfl_xor_sync(FULL, pk[0], 1);
            if (!(lane & 1)) {
                const float gv = dinv[s] * pk[0];
                G[(size_t)s * H + (lane >> 1)] = gv;
                mirror_store(mirror, (size_t)s * H + (lane >> 1), gv);
            }
        }
        }   // labelled rows of the block
        if (terms && my_truth >= 0) terms[my_slot] = my_term;
    }

    __shared__ float s_loss[L2_WARPS];
    __shared__ int s_count[L2_WARPS], s_wrong[L2_WARPS];
    if (lane == 0) { s_loss[warp] = loss; s_count[warp] = count; s_wrong[warp] = wrong; }
    if (training) {
#pragma unroll
        for (int t = 0; t < CPL; t++) {
            const int cls = lane + 32 * t;
            if (cls < c)
#pragma unroll
                for (int k = 0; k < H; k++) sAcc[warp * H * c + k * c + cls] = dw[t][k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f; int cn = 0, wr = 0;
        for (int w = 0; w < L2_WARPS; w++) { l += s_loss[w]; cn += s_count[w]; wr += s_wrong[w]; }
        ce_partials[blockIdx.x] = L2Partial{l, cn, wr, 0};
    }
    if (training) {
        float *out = dw_partials + (size_t)blockIdx.x * H * c;
        for (int i = threadIdx.x; i < H * c; i += L2_THREADS) {
            float v = 0.f;
            for (int w = 0; w < L2_WARPS; w++) v += sAcc[w * H * c + i];
            out[i] = v;
        }
    }
}

// the last block produces the scalar result; the others reduce 32 elements of dW2 each over the CTA partials
__global__ void __launch_bounds__(256) layer2_finish_kernel(const L2Partial *__restrict__ ce_partials, const float *__restrict__ dw_partials,
                                                             int parts, int hc, int training, float *__restrict__ W2_grad,
                                                             gcnk_ce_result *__restrict__ result, float *__restrict__ red4) {
    if (blockIdx.x + 1 < gridDim.x) {
        reduce_parts_block(dw_partials, W2_grad, hc, parts);
        return;
    }
    __shared__ float s_loss[256];
    __shared__ int s_count[256], s_wrong[256];
    float l = 0.f; int cn = 0, wr = 0;
    for (int b = threadIdx.x; b < parts; b += 256) { l += ce_partials[b].loss; cn += ce_partials[b].count; wr += ce_partials[b].wrong; }
    s_loss[threadIdx.x] = l; s_count[threadIdx.x] = cn; s_wrong[threadIdx.x] = wr;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
            s_count[threadIdx.x] += s_count[threadIdx.x + o];
            s_wrong[threadIdx.x] += s_wrong[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        result->loss = s_loss[0] / (float)s_count[0];
        result->count = s_count[0]; result->wrong = s_wrong[0]; result->pad = 0;
        // raw sums as floats (counts < 2^24 are exact): what a row-partitioned run all-reduces across ranks
        red4[0] = s_loss[0]; red4[1] = (float)s_count[0]; red4[2] = (float)s_wrong[0]; red4[3] = 0.f;
    }
}

// one 32-row block per warp when that gives at most ~6 waves of CTAs, else a multiple of the SM count
int l2_grid(int n) {
    const int blocks = (n + 31) / 32, ctas = (blocks + L2_WARPS - 1) / L2_WARPS;
    return std::max(1, std::min(ctas, sm_count() * 12));
}

}  // namespace

extern "C" {

size_t gcnk_layer2_workspace(int n, int h, int c) {
    const size_t g = (size_t)l2_grid(n);
    return 16 + g * sizeof(L2Partial) + g * (size_t)h * c * sizeof(float);
}

int gcnk_layer2_fused(const float *P, const float *W2, const int *split, const int *label, int current_split, int n, int h,
                      int c, int training, int count, const float *d_dinv, float *G_scaled, float *W2_grad, float *logits_out,
                      gcnk_ce_result *d_result, float *workspace, size_t workspace_bytes, gcnk_stream_t stream) {
    return gcnk_layer2_fused_terms(P, W2, split, label, current_split, n, h, c, training, count, d_dinv, G_scaled, W2_grad, logits_out,
                                   d_result, workspace, workspace_bytes, nullptr, nullptr, stream);
}

int gcnk_layer2_fused_terms(const float *P, const float *W2, const int *split, const int *label, int current_split, int n, int h,
                            int c, int training, int count, const float *d_dinv, float *G_scaled, float *W2_grad, float *logits_out,
                            gcnk_ce_result *d_result, float *workspace, size_t workspace_bytes, float *loss_terms,
                            const int *term_index, gcnk_stream_t stream) {
    GCNK_REQUIRE(P && W2 && split && label && d_result && n >= 0 && h > 0 && c > 0, "bad arguments");
    GCNK_REQUIRE(c <= 32 * CPL_MAX && (size_t)h * c <= 4096, "needs c <= 128 and h*c <= 4096");
    GCNK_REQUIRE(!training || (G_scaled && W2_grad && d_dinv), "training needs G, W2_grad and dinv");
    GCNK_REQUIRE(workspace && workspace_bytes >= gcnk_layer2_workspace(n, h, c), "workspace too small");
    cudaStream_t st = S(stream);
    const int grid = l2_grid(n);
    float *red4 = workspace;                                               // {sum of loss terms, count, wrong, 0}
    L2Partial *ce_partials = reinterpret_cast<L2Partial *>(workspace + 4);
    float *dw_partials = reinterpret_cast<float *>(ce_partials + grid);
    const size_t smem = sizeof(float) * ((size_t)h * c + L2_WARPS * (size_t)c + L2_WARPS * (size_t)h +
                                         (training ? L2_WARPS * (size_t)h * c : 0));
    const Mirror mirror = training ? take_mirror(G_scaled) : Mirror{};
    if (h == 16 && c <= 64) {
        // registers hold the per-lane slice of dW2; the shared accumulator is only the end-of-kernel exchange
        if (c <= 32)
            layer2_h16_kernel<1><<<grid, L2_THREADS, smem, st>>>(P, W2, split, label, current_split, n, c, training, (float)count,
                                                                d_dinv, G_scaled, logits_out, ce_partials, dw_partials, mirror, loss_terms, term_index);
        else
            layer2_h16_kernel<2><<<grid, L2_THREADS, smem, st>>>(P, W2, split, label, current_split, n, c, training, (float)count,
                                                                d_dinv, G_scaled, logits_out, ce_partials, dw_partials, mirror, loss_terms, term_index);
        GCNK_LAUNCHED();
    } else {
        GCNK_REQUIRE(mirror.n == 0, "mirrored output is only implemented for hidden width 16");
        if (smem > 48 * 1024) {
            GCNK_REQUIRE(smem <= 200 * 1024, "h*c too large for shared memory");
            GCNK_CUDA(cudaFuncSetAttribute(layer2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        layer2_kernel<<<grid, L2_THREADS, smem, st>>>(P, W2, split, label, current_split, n, h, c, training, (float)count, d_dinv,
                                                      G_scaled, logits_out, ce_partials, dw_partials, loss_terms, term_index);
        GCNK_LAUNCHED();
    }
    const int hc = h * c;
    layer2_finish_kernel<<<(training ? (hc + 31) / 32 : 0) + 1, 256, 0, st>>>(ce_partials, dw_partials, grid, hc, training, W2_grad, d_result, red4);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // extern "C"
