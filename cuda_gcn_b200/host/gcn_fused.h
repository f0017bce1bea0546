// gcn_fused.h — private state of the fused plans (gcn.cpp: hidden 16; gcn_wide.cpp: wide hidden).  Not part of the
// reference-shaped public headers.
#pragma once
#include <cstdint>

#include <vector>

#include "check.h"
#include "gcn.h"

// The reference's scalar loss loop is reproduced bit for bit (gcnk_sequential_sum) for splits of at least this many labelled
// rows; below it the parallel sum is used: with fewer terms the scalar loop's own rounding stays far inside the parity
// tolerance (measured at Reddit shape, 23,000 validation rows: <= 5.4e-6 relative over 30 epochs, against 1.0e-4 for the
// 153,756 training rows), and the sum would sit on the critical path of every eval pass.
constexpr int SEQ_LOSS_MIN_ROWS = 65536;

template <typename T>
static inline T *upload(const std::vector<T> &h) {
    T *d = nullptr;
    GCNK_CHECK(gcnk_malloc((void **)&d, sizeof(T) * h.size()));
    GCNK_CHECK(gcnk_memcpy_h2d(d, h.data(), sizeof(T) * h.size(), nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
    return d;
}

// Buffers of the fused plan.  "_s" = already multiplied by d^-1/2 of its own row (the gather kernels
// take pre-scaled sources, so an edge costs one index and one row read; see csrc/graph.cu).
struct GCN::Fused {
    // ---- wide plan (gcn_wide.cpp): hidden*classes too large for the row-local layer-2 kernel
    bool wide = false, wide_sources_owned = false;
    int Cp = 0;                // classes rounded up to 4: the pitch of every class-width buffer (zero padding columns)
    size_t buf_floats = 0;     // floats per exchanged gather source (N*H, wide: N*Cp)
    float *T_s = nullptr;      // [N x Cp]   dinv (.) (H1 * W2): source of the forward class-width GraphSum (exchanged)
    float *D_s = nullptr;      // [N x Cp]   dinv (.) dlogits: source of the backward one (exchanged)
    float *X_all = nullptr;    // [N x F]    pristine features of ALL nodes (the layer-1 gather reads every node's row)
    bool x_all_owned = false;
    float *X_all_spare = nullptr;   // row-partitioned: the buffer the next input (all ranks' rows) is assembled in (epoch_prefetch)
    bool spare_needs_gather = false;
    float *Xd_s = nullptr;     // [N x F]    dinv (.) dropout(X): the layer-1 gather source of this pass
    float *AXd = nullptr;      // [n x F]    A_hat * dropout(X)   (forward input of X W1 AND the left factor of dW1)
    float *AXw = nullptr;      // [n x F]    A_hat * X, static: eval passes need no gather at all in layer 1
    float *H1 = nullptr;       // [n x H]    Z1 = AXd*W1, then dropout(relu(Z1)) in place
    float *Tl = nullptr;       // [n x Cp]   H1 * W2 (local rows, before the pre-scale into T_s)
    float *logits = nullptr;   // [n x Cp]   A_hat * T (labelled rows of the split only)
    float *dT = nullptr;       // [n x Cp]   A_hat * dlogits
    float *dH1 = nullptr;      // [n x H]    dT * W2^T, then masked in place = dZ1
    float *W2p = nullptr, *dW2p = nullptr;   // [H x Cp] padded copies of W2 / its gradient
    uint32_t *wkeep0 = nullptr, *wkeep1 = nullptr, *wmask = nullptr;   // keep bits of X (all N*F), of H1 (own rows), ReLU&dropout mask
    float *mm_ws = nullptr; size_t mm_ws_bytes = 0;                     // split-K workspace of the two weight-gradient GEMMs
    const float *dinv_all = nullptr;                                    // [N] d^-1/2 of every node

    float *xw_s = nullptr;     // [N x H]  dinv (.) (dropout(X) * W1)
    float *h1_s = nullptr;     // [N x H]  dinv (.) dropout(relu(A_hat * X W1))
    float *P = nullptr;        // [N x H]  A_hat * H1
    float *G = nullptr;        // [N x H]  dinv (.) (dlogits * W2^T)
    float *Gm = nullptr;       // [N x H]  dinv (.) dropout'/relu'(A_hat * dlogits W2^T)
    float *dxw = nullptr;      // [N x H]  gradient wrt X W1
    uint32_t *keep0 = nullptr, *keep1 = nullptr, *mask = nullptr;   // the keep bits this pass reads
    // Double-buffered keep bits: the masks of the NEXT training pass are drawn on a side stream while this pass
    // (and the eval pass after it) runs — the generator is ALU-bound, the gathers are L2-bound.  The draws are a
    // pure function of the stream position, so the bits are identical to drawing them in line; if anything else
    // consumed the shared stream in between (state mismatch) they are simply drawn again in line.
    uint32_t *keep0_buf[2] = {nullptr, nullptr}, *keep1_buf[2] = {nullptr, nullptr};
    int cur = 0;
    gcnk_stream_t rng_stream = nullptr;
    void *ev_ready = nullptr, *ev_go = nullptr;
    bool pre_valid = false;
    uint64_t pre_state[2] = {0, 0};
    float *ws = nullptr; size_t ws_bytes = 0;
    gcnk_ce_result *d_result = nullptr, *h_result = nullptr;   // device / pinned host
    float *d_sumsq = nullptr, *h_sumsq = nullptr;
    float *h_red = nullptr;    // pinned [2][4]: {sum of loss terms, count, wrong, 0} after the cross-rank reduction, per result slot
    float sumsq_used[2] = {0.f, 0.f};
    bool seq_used[2] = {false, false};
    bool sumsq_pending = false;
    gcnk_rng *slice_rng = nullptr;   // positions a copy of the shared stream at this rank's rows
    float sumsq = 0;           // sum(W1^2) of the current weights
    // Views of the graph for the passes that need only part of A_hat*x (splits are static, so these are built once):
    //   rows[s]     only the labelled rows of split s are aggregated — the loss, the accuracy and the layer-2
    //               gradients never look at the logits of any other row (module.cpp:130-133: truth < 0 rows are skipped)
    //   cols_train  entries pointing at rows outside the training split dropped — their loss gradient is exactly zero
    gcnk_graph *rows[4] = {nullptr, nullptr, nullptr, nullptr}, *cols_train = nullptr;
    int *keep[4] = {nullptr, nullptr, nullptr, nullptr};
    // AX = A_hat * X, computed once when X is dense: without input dropout (every eval pass)
    // A_hat*(X*W1) = (A_hat*X)*W1 is one streaming pass and no gather
    float *AX = nullptr;
    bool ax_valid = false, use_views = true;
    // TMA path of the dense feature transform: packed copies (row pitch ld floats, a multiple of 32) of X and A_hat*X
    float *Xp = nullptr, *AXp = nullptr, *bw_ws = nullptr;
    size_t bw_ws_bytes = 0;
    int ld = 0;
    bool xp_dirty = false, tc_transform = false;
    // Row-partitioned runs: the four gather sources live in ONE slab that every peer maps over NVLink (CUDA IPC);
    // producers mirror their rows into the peers' slabs and a flag barrier replaces the all-gather collective.
    bool p2p = false;
    float *slab = nullptr;
    void *peer_slab[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int *flag_arrays[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int barrier_value = 0, world = 1, rank = 0;
    // Exchange by push + signal: buffer b of {xw_s, h1_s, G, Gm} has one flag per producing rank in every rank's slab
    // (ints [64 + 8 b + r] of the flag block); seq[b] counts how often the buffer has been produced, and is the value
    // the pushes publish and the consuming gather waits for.  halo[p] (optional) lists the local rows peer p references.
    int seq[5] = {0, 0, 0, 0, 0};          // [4]: the loss-term buffer
    // Reference-order loss: layer 2 stores every labelled row's loss term at its rank among the labelled rows of the split
    // (term_index[split][row], global order), and one warp adds them up exactly as the reference's scalar loop does
    // (gcnk_sequential_sum) on a side stream, under the backward pass.  Row-partitioned: every rank pushes its compact
    // range to the peers and every rank computes the same global sum.  GCN_TREE_LOSS=1: the plain parallel sum instead.
    bool seq_loss = true;
    int seq_when = 1;          // training pass: 0 = right behind layer 2 (under the backward gathers), 1 = behind the last gather (under the
                               // weight-gradient kernel), 2 = on the main stream at the end of the pass.  GCN_SEQ_WHEN
    bool seq_joined = true;    // the main stream has waited for the side stream's last loss sum
    int seq_pending = 0;       // split whose loss sum is still to be launched in this pass
    float *terms = nullptr, *d_seq = nullptr, *h_seq = nullptr;
    bool terms_owned = false;
    int *term_index[4] = {nullptr, nullptr, nullptr, nullptr};
    int term_c0[4] = {0, 0, 0, 0}, term_cnt[4] = {0, 0, 0, 0}, term_len[4] = {0, 0, 0, 0};
    size_t term_region = 0;     // floats per split region of `terms`
    gcnk_stream_t seq_stream = nullptr;
    void *ev_l2 = nullptr, *ev_seq = nullptr;
    int *halo_rows[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int halo_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool use_halo = false, signal_exchange = true;
    // Wide plan: what a peer reads of an exchanged class-width source is much less than all rows — the forward GraphSum only
    // aggregates the labelled rows of the split (their neighbours), the backward one only reads training columns.  Per
    // consumer a row list per peer (gcn_wide.cpp: build_wide_halo); `halo_now` selects the set the next publish() uses.
    struct HaloSet { int *rows[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; int count[8] = {0, 0, 0, 0, 0, 0, 0, 0}; bool valid = false; };
    HaloSet halo_T[4], halo_D;
    const HaloSet *halo_now = nullptr;
    gcnk_stream_t stream = nullptr;   // everything the fused plan enqueues runs on this (non-blocking) stream
    // Overlap of the exchange with the local part of the consuming GraphSum (GCN::exchange_overlapped; GCN_OVERLAP=0 turns
    // it off): own-columns / remote-columns views of the CSR slice (gt_*: training columns only; rows_*: per split), the
    // raw partial sums, and the communication stream the pushes run on.
    bool overlap = false, fused_xchg = false;
    gcnk_graph *g_own = nullptr, *g_rem = nullptr, *gt_own = nullptr, *gt_rem = nullptr;
    gcnk_graph *rows_own[4] = {nullptr, nullptr, nullptr, nullptr}, *rows_rem[4] = {nullptr, nullptr, nullptr, nullptr};
    float *partial = nullptr;
    gcnk_stream_t comm_stream = nullptr;
    void *ev_prod = nullptr;
    unsigned *d_counter2 = nullptr;
    int *d_err = nullptr, *h_err = nullptr;
    int *h_async = nullptr;      // pinned copy of the kernel library's async error flag (mbarrier time-outs)
    unsigned *d_counter = nullptr;
    float *areas[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // all-reduce exchange areas
    size_t slot_floats = 0;
    gcnk_comm *comm = nullptr;
    ~Fused() {
        gcnk_device_sync();
        if (slab) {
            // nobody may still be writing into this slab (or reading ours) when it goes away
            if (comm) { float *b[1] = {(float *)slab}; const size_t c1[1] = {1}; gcnk_comm_allreduce(comm, b, c1, 1, 1, nullptr); gcnk_device_sync(); }
            for (int r = 0; r < world; r++) if (r != rank && peer_slab[r]) gcnk_ipc_release(peer_slab[r]);
            gcnk_free(slab);
            xw_s = h1_s = G = Gm = nullptr;                 // they were views into the slab
        }
        if (d_err) gcnk_free(d_err);
        if (d_counter) gcnk_free(d_counter);
        if (h_err) gcnk_free_host(h_err);
        if (h_async) gcnk_free_host(h_async);
        if (rng_stream) { gcnk_stream_sync(rng_stream); gcnk_stream_destroy(rng_stream); }
        if (comm_stream) { gcnk_stream_sync(comm_stream); gcnk_stream_destroy(comm_stream); }
        if (ev_prod) gcnk_event_destroy(ev_prod);
        for (gcnk_graph *v : {rows_own[1], rows_own[2], rows_own[3], rows_rem[1], rows_rem[2], rows_rem[3]}) if (v) gcnk_graph_destroy(v);   // row views first: they borrow
        for (gcnk_graph *v : {g_own, g_rem, gt_own, gt_rem}) if (v) gcnk_graph_destroy(v);
        if (partial) gcnk_free(partial);
        if (d_counter2) gcnk_free(d_counter2);
        if (seq_stream) { gcnk_stream_sync(seq_stream); gcnk_stream_destroy(seq_stream); }
        if (stream) { gcnk_stream_sync(stream); gcnk_stream_destroy(stream); }
        if (ev_l2) gcnk_event_destroy(ev_l2);
        if (ev_seq) gcnk_event_destroy(ev_seq);
        if (terms_owned && terms) gcnk_free(terms);
        if (d_seq) gcnk_free(d_seq);
        if (h_seq) gcnk_free_host(h_seq);
        for (int *t : term_index) if (t) gcnk_free(t);
        if (wide_sources_owned && T_s) gcnk_free(T_s);
        if (x_all_owned && X_all) gcnk_free(X_all);
        if (X_all_spare) gcnk_free(X_all_spare);
        for (void *q : {(void *)Xd_s, (void *)AXd, (void *)AXw, (void *)H1, (void *)Tl, (void *)logits, (void *)dT, (void *)dH1, (void *)W2p,
                        (void *)dW2p, (void *)wkeep0, (void *)wkeep1, (void *)wmask, (void *)mm_ws})
            if (q) gcnk_free(q);
        for (int *h : halo_rows) if (h) gcnk_free(h);
        for (HaloSet *hs : {&halo_T[1], &halo_T[2], &halo_T[3], &halo_D}) for (int *h : hs->rows) if (h) gcnk_free(h);
        if (ev_ready) gcnk_event_destroy(ev_ready);
        if (ev_go) gcnk_event_destroy(ev_go);
        for (uint32_t *b : {keep0_buf[1], keep1_buf[1]}) if (b) gcnk_free(b);
        for (gcnk_graph *g : {rows[1], rows[2], rows[3], cols_train}) if (g) gcnk_graph_destroy(g);
        for (int *k : keep) if (k) gcnk_free(k);   // keep[0] = global train-column flags
        if (AX) gcnk_free(AX);
        for (float *b : {Xp, AXp, bw_ws}) if (b) gcnk_free(b);
        for (void *p : {(void *)xw_s, (void *)h1_s, (void *)P, (void *)G, (void *)Gm, (void *)dxw, (void *)keep0_buf[0], (void *)keep1_buf[0],
                        (void *)mask, (void *)ws, (void *)d_result, (void *)d_sumsq})
            if (p) gcnk_free(p);
        if (h_result) gcnk_free_host(h_result);
        if (h_sumsq) gcnk_free_host(h_sumsq);
        if (h_red) gcnk_free_host(h_red);
        if (slice_rng) gcnk_rng_destroy(slice_rng);
    }
};

