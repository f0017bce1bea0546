// gcn_wide.cpp — the fused plan for WIDE hidden layers (BASELINE configs[4]: ogbn-products shape, 100 dense features ->
// hidden 256 -> 47 classes), single-GPU and row-partitioned.  Same model and the same numbers as the reference's module
// chain (src/seq/gcn.cpp:20-65, module.cpp) within fp32 rounding; what changes is the order of the products:
//
//   layer 1   A_hat * (drop(X) * W1)  ->  (A_hat * drop(X)) * W1      the GraphSum runs at the INPUT width (100) instead of
//             the hidden width (256), and AXd = A_hat*drop(X) is kept: dW1 = drop(X)^T * (A_hat^T dZ1) = AXd^T * dZ1 for the
//             symmetric A_hat, so the backward needs no hidden-width GraphSum either.  Eval passes (no dropout) use the
//             static A_hat*X computed once: no layer-1 gather at all.
//   layer 2   A_hat * (H1 * W2): the reference's order, at the class width (47, stored with pitch 48).  Only the labelled
//             rows of the split are aggregated forward; the backward A_hat*dlogits reads only training columns.
//
// Row-partitioned: every rank holds ALL rows of X (static, 980 MB at products shape) and draws all N*F keep bits from
// the shared xorshift128+ stream, so layer 1 needs NO exchange; per training pass only the two class-width sources
// (T_s = dinv.(H1 W2) and D_s = dinv.dlogits, 2 x N x 48 floats) travel, by the same push + arrival-flag scheme as the
// hidden-16 plan (gcn.cpp: publish/await), and the weight gradients are summed by the peer-memory all-reduce.
// The GEMMs go to the tensor cores (csrc/matmul_tc.cu: tcgen05, TMEM, TMA) where the shape allows.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "check.h"
#include "gcn_fused.h"
#include "rand.h"
#include "timer.h"

void GCN::build_wide() {
    Fused &z = *fz;
    const int N = params.num_nodes, F = params.input_dim, H = params.hidden_dim, Cp = z.Cp;
    gcnk_graph *g = graph_handle();
    const float *dinv_loc = nullptr;
    GCNK_CHECK(gcnk_graph_dinv(g, &dinv_loc));
    z.dinv_all = dist.world > 1 ? d_dinv_global : dinv_loc;

    // every node's features: the layer-1 gather source is built from them every training pass
    if (dist.world > 1) { z.X_all = upload(full_data->feature_value); z.x_all_owned = true; }
    else z.X_all = d_feature_value;

    const size_t nf_all = sizeof(float) * (size_t)N * F, nf_loc = sizeof(float) * (size_t)n_loc * F;
    const size_t nh_loc = sizeof(float) * (size_t)n_loc * H, nc_loc = sizeof(float) * (size_t)n_loc * Cp;
    GCNK_CHECK(gcnk_malloc((void **)&z.Xd_s, nf_all));
    GCNK_CHECK(gcnk_malloc((void **)&z.AXd, nf_loc));
    GCNK_CHECK(gcnk_malloc((void **)&z.AXw, nf_loc));
    GCNK_CHECK(gcnk_malloc((void **)&z.H1, nh_loc));
    GCNK_CHECK(gcnk_malloc((void **)&z.dH1, nh_loc));
    for (float **p : {&z.Tl, &z.logits, &z.dT}) GCNK_CHECK(gcnk_malloc((void **)p, nc_loc));
    GCNK_CHECK(gcnk_memset(z.logits, 0, nc_loc, nullptr));
    for (float **p : {&z.W2p, &z.dW2p}) GCNK_CHECK(gcnk_malloc((void **)p, sizeof(float) * (size_t)H * Cp));
    // keep bits, double-buffered: the NEXT training pass's draws run on a low-priority side stream under this pass's GEMMs
    // (the generator is ALU-bound, the tcgen05 kernels leave the ALUs idle) — same scheme as the hidden-16 plan
    for (int b = 0; b < 2; b++) {
        GCNK_CHECK(gcnk_malloc((void **)&z.keep0_buf[b], sizeof(uint32_t) * ((size_t)N * F / 32 + 4)));
        GCNK_CHECK(gcnk_malloc((void **)&z.keep1_buf[b], sizeof(uint32_t) * ((size_t)n_loc * H / 32 + 4)));
    }
    z.keep0 = z.keep0_buf[0]; z.keep1 = z.keep1_buf[0];
    {
        const char *ns = getenv("GCN_NO_RNG_OVERLAP");
        if (!(ns && *ns && strcmp(ns, "0"))) {
            GCNK_CHECK(gcnk_stream_create_low_priority(&z.rng_stream));
            GCNK_CHECK(gcnk_event_create(&z.ev_ready));
            GCNK_CHECK(gcnk_event_create(&z.ev_go));
        }
    }
    GCNK_CHECK(gcnk_malloc((void **)&z.wmask, sizeof(uint32_t) * ((size_t)n_loc * H / 32 + 4)));
    z.mm_ws_bytes = std::max(gcnk_matmul_tn_workspace(n_loc, F, H), gcnk_matmul_tn_workspace(n_loc, H, Cp));
    GCNK_CHECK(gcnk_malloc((void **)&z.mm_ws, std::max<size_t>(z.mm_ws_bytes, 16)));

    // views: labelled rows per split (forward class-width GraphSum), training columns (backward)
    for (int s = 1; s <= 3; s++) {
        std::vector<int> keep((size_t)n_loc);
        for (int i = 0; i < n_loc; i++) keep[i] = data->split[i] == s && data->label[i] >= 0;
        z.keep[s] = upload(keep);
        if (s < 3) GCNK_CHECK(gcnk_graph_create_view(&z.rows[s], g, z.keep[s], nullptr, nullptr));   // test split: on first use
    }
    std::vector<int> train_cols((size_t)N);
    for (int i = 0; i < N; i++) train_cols[i] = full_data->split[i] == 1 && full_data->label[i] >= 0;
    z.keep[0] = upload(train_cols);
    GCNK_CHECK(gcnk_graph_create_view(&z.cols_train, g, nullptr, z.keep[0], nullptr));

    if (dist.world > 1 && z.p2p && z.signal_exchange) build_wide_halo();

    // A_hat * X for this rank's rows, once (eval passes)
    GCNK_CHECK(gcnk_drop_scale_rows(z.X_all, N, F, nullptr, 1.0f, z.dinv_all, z.Xd_s, nullptr));
    GCNK_CHECK(gcnk_gather_plain(g, z.Xd_s, z.AXw, F, nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
    z.ax_valid = true;
}

// Halo lists of the wide plan.  Rank p reads, of the forward source T_s, only the columns its labelled rows of
// the current split are adjacent to (at products shape: 8 % training rows x ~50 neighbours = about 40 % of all nodes, 10 %
// for the validation split), and of the backward source D_s only training columns it is adjacent to (8 % of all nodes).
// Every rank marks those columns in bit maps (one per split + one for "any row"), the maps are all-gathered once, and
// each rank keeps, per consumer and peer, the list of its own rows to send.  GCN_HALO=0 turns the lists off.
void GCN::build_wide_halo() {
    Fused &z = *fz;
    const char *hv = getenv("GCN_HALO");
    if (hv && *hv && !strcmp(hv, "0")) return;
    const int N = params.num_nodes;
    const size_t bytes = ((size_t)N + 7) / 8;
    std::vector<unsigned char> mine(4 * bytes, 0), all(4 * bytes * (size_t)dist.world, 0);
    const std::vector<int> &ip = data->graph.indptr, &ix = data->graph.indices;
    for (int i = 0; i < n_loc; i++) {
        const int sp = data->label[i] >= 0 && data->split[i] >= 1 && data->split[i] <= 3 ? data->split[i] : 0;
        for (int e = ip[i]; e < ip[i + 1]; e++) {
            const int c = ix[e];
            mine[(size_t)c >> 3] |= (unsigned char)(1u << (c & 7));                              // map 0: any row
            if (sp) mine[(size_t)sp * bytes + ((size_t)c >> 3)] |= (unsigned char)(1u << (c & 7));
        }
    }
    GCNK_CHECK(gcnk_comm_allgather_bytes(dist.comm, mine.data(), all.data(), (int)(4 * bytes)));
    for (int p = 0; p < dist.world; p++) {
        if (p == dist.rank) continue;
        const unsigned char *maps = all.data() + 4 * bytes * (size_t)p;
        auto bit = [&](int map, int i) { return (maps[(size_t)map * bytes + ((size_t)i >> 3)] >> (i & 7)) & 1u; };
        for (int sp = 1; sp <= 3; sp++) {
            std::vector<int> rows;
            for (int i = r0; i < r0 + n_loc; i++) if (bit(sp, i)) rows.push_back(i - r0);
            z.halo_T[sp].count[p] = (int)rows.size();
            if (!rows.empty()) z.halo_T[sp].rows[p] = upload(rows);
            z.halo_T[sp].valid = true;
        }
        std::vector<int> rows;
        for (int i = r0; i < r0 + n_loc; i++)
            if (bit(0, i) && data->split[i - r0] == 1 && data->label[i - r0] >= 0) rows.push_back(i - r0);
        z.halo_D.count[p] = (int)rows.size();
        if (!rows.empty()) z.halo_D.rows[p] = upload(rows);
        z.halo_D.valid = true;
    }
}

void GCN::wide_enqueue(int current_split, bool training, int slot) {
    consume_pending_input();
    Fused &z = *fz;
    gcnk_stream_t st = z.stream;
    gpu_timer_set_stream(st);
    const int N = params.num_nodes, F = params.input_dim, H = params.hidden_dim, C = params.output_dim, Cp = z.Cp;
    const float p = params.dropout;
    const float scale = 1 / (1 - p);                                      // module.cpp:212
    const bool drop = training && (int)(p * (float)MY_RAND_MAX) > 0;      // threshold 0 keeps everything
    gcnk_graph *g = graph_handle();
    const float *dinv = nullptr;                                          // d^-1/2 of the local rows
    GCNK_CHECK(gcnk_graph_dinv(g, &dinv));
    Variable &W1 = variables[2], &W2 = variables[5];
    const size_t own = (size_t)r0 * Cp;                                   // this rank's slice of an [N x Cp] gather source
    const int sidx = current_split >= 1 && current_split <= 3 ? current_split : 0;
    gcnk_graph *g_rows = g;
    if (sidx) {
        if (!z.rows[sidx]) {
            GCNK_CHECK(gcnk_graph_create_view(&z.rows[sidx], g, z.keep[sidx], nullptr, nullptr));
            GCNK_CHECK(gcnk_stream_sync(nullptr));
        }
        g_rows = z.rows[sidx];
    }

    const float *ax = z.AXw;                                              // layer-1 input: A_hat * X (eval) or A_hat * drop(X)
    if (!training && !z.ax_valid) {                                       // the features were replaced (set_input_from_host)
        GCNK_CHECK(gcnk_drop_scale_rows(z.X_all, N, F, nullptr, 1.0f, z.dinv_all, z.Xd_s, st));
        GCNK_CHECK(gcnk_gather_plain(g, z.Xd_s, z.AXw, F, st));
        z.ax_valid = true;
    }
    if (training) {
        // The reference draws N*F values (input dropout) and then N*H values (hidden dropout) per training pass from ONE
        // stream, in element order (module.cpp:214-218 via gcn.cpp:110-111).  Every rank needs the input bits of ALL
        // nodes (its gather reads every node's row) and the hidden bits of its own rows.
        gpu_timer_begin(TMR_DROPOUT_FW);
        uint64_t state[2];
        GCNK_CHECK(gcnk_rng_get_state(global_rng(), state));
        auto draw_masks = [&](const uint64_t *from, uint32_t *k0, uint32_t *k1, gcnk_stream_t stream) {
            GCNK_CHECK(gcnk_rng_set_state(z.slice_rng, from[0], from[1]));
            GCNK_CHECK(gcnk_dropout_mask(z.slice_rng, k0, (int64_t)N * F, p, stream));
            GCNK_CHECK(gcnk_rng_set_state(z.slice_rng, from[0], from[1]));
            GCNK_CHECK(gcnk_rng_skip(z.slice_rng, (uint64_t)N * F + (uint64_t)r0 * H));
            GCNK_CHECK(gcnk_dropout_mask(z.slice_rng, k1, (int64_t)n_loc * H, p, stream));
        };
        if (drop) {
            z.keep0 = z.keep0_buf[z.cur]; z.keep1 = z.keep1_buf[z.cur];
            const bool ahead = z.pre_valid && z.pre_state[0] == state[0] && z.pre_state[1] == state[1];
            if (z.pre_valid) GCNK_CHECK(gcnk_stream_wait_event(st, z.ev_ready));      // the side stream's last draw targets this buffer pair
            if (!ahead) draw_masks(state, z.keep0, z.keep1, st);
            z.pre_valid = false;
        }
        GCNK_CHECK(gcnk_rng_skip(global_rng(), (uint64_t)N * F + (uint64_t)N * H));   // consumed even when p == 0
        gpu_timer_end(TMR_DROPOUT_FW);

        // M0 Dropout + M2 GraphSum, re-ordered in front of M1: AXd = A_hat * drop(X) for the local rows
        gpu_timer_begin(TMR_GRAPHSUM_FW);
        GCNK_CHECK(gcnk_drop_scale_rows(z.X_all, N, F, drop ? z.keep0 : nullptr, scale, z.dinv_all, z.Xd_s, st));
        GCNK_CHECK(gcnk_gather_plain(g, z.Xd_s, z.AXd, F, st));
        if (drop && z.rng_stream) {
            // the next training pass's bits, into the other buffer pair, released now: the rest of this pass is GEMMs and
            // narrow gathers (the other buffers' last readers finished with the previous training pass: host sync)
            GCNK_CHECK(gcnk_event_record(z.ev_go, st));
            GCNK_CHECK(gcnk_stream_wait_event(z.rng_stream, z.ev_go));
            GCNK_CHECK(gcnk_rng_get_state(global_rng(), z.pre_state));
            draw_masks(z.pre_state, z.keep0_buf[z.cur ^ 1], z.keep1_buf[z.cur ^ 1], z.rng_stream);
            GCNK_CHECK(gcnk_event_record(z.ev_ready, z.rng_stream));
            z.pre_valid = true;
            z.cur ^= 1;
        }
        gpu_timer_end(TMR_GRAPHSUM_FW);
        ax = z.AXd;
    }
    // M1 (Sparse)Matmul: Z1 = AX * W1; M3 ReLU + M4 Dropout in place
    gpu_timer_begin(TMR_SPMATMUL_FW);
    GCNK_CHECK(gcnk_matmul_nn(ax, F, W1.data, H, z.H1, H, n_loc, F, H, nullptr, st));
    GCNK_CHECK(gcnk_relu_dropout_fw(z.H1, (int64_t)n_loc * H, drop ? z.keep1 : nullptr, training ? scale : 1.0f, training ? z.wmask : nullptr, st));
    gpu_timer_end(TMR_SPMATMUL_FW);
    // M5 Matmul: T = H1 * W2 (padded to Cp columns), pre-scaled by d^-1/2 straight into the exchanged gather source
    gpu_timer_begin(TMR_MATMUL_FW);
    GCNK_CHECK(gcnk_pad_cols(W2.data, z.W2p, H, C, Cp, st));
    GCNK_CHECK(gcnk_matmul_nn(z.H1, H, z.W2p, Cp, z.T_s + own, Cp, n_loc, H, Cp, dinv, st));
    gpu_timer_end(TMR_MATMUL_FW);
    if (sidx && z.halo_T[sidx].valid) z.halo_now = &z.halo_T[sidx];
    publish(z.T_s, Cp);
    // M6 GraphSum at the class width, only for the rows whose logits the loss looks at
    gpu_timer_begin(TMR_GATHER_PART);
    await(z.T_s, Cp);
    GCNK_CHECK(gcnk_gather_plain(g_rows, z.T_s, z.logits, Cp, st));
    gpu_timer_end(TMR_GATHER_PART);

    // M7 CrossEntropyLoss + get_accuracy; count = labelled rows of the split over ALL ranks (module.cpp:154-158)
    gpu_timer_begin(TMR_LOSS_FW);
    const bool seq = z.seq_loss && sidx != 0 && split_count[sidx] >= SEQ_LOSS_MIN_ROWS;
    GCNK_CHECK(gcnk_ce_rows(z.logits, Cp, d_split, d_label, current_split, n_loc, C, training, split_count[current_split & 3], dinv,
                            training ? z.D_s + own : nullptr, z.d_result, z.ws, z.ws_bytes, seq ? z.terms + (size_t)(sidx - 1) * z.term_region : nullptr,
                            seq ? z.term_index[sidx] : nullptr, st));
    gpu_timer_end(TMR_LOSS_FW);
    if (seq) enqueue_loss_sum(sidx, training, slot);
    z.seq_used[slot] = seq;

    if (training) {
        // M6 backward: dT = A_hat * dlogits (all rows; only training columns carry a gradient)
        if (z.halo_D.valid) z.halo_now = &z.halo_D;
        publish(z.D_s, Cp);
        gpu_timer_begin(TMR_GRAPHSUM_BW);
        await(z.D_s, Cp);
        GCNK_CHECK(gcnk_gather_plain(z.cols_train, z.D_s, z.dT, Cp, st));
        gpu_timer_end(TMR_GRAPHSUM_BW);
        if (z.seq_when == 1) flush_loss_sum();                           // under the backward GEMMs
        // M5 backward: dW2 = H1^T dT, dH1 = dT W2^T;  M4/M3 backward: the mask;  M1 backward: dW1 = AXd^T dZ1
        gpu_timer_begin(TMR_MATMUL_BW);
        GCNK_CHECK(gcnk_matmul_tn(z.H1, H, z.dT, Cp, z.dW2p, Cp, n_loc, H, Cp, z.mm_ws, z.mm_ws_bytes, st));
        GCNK_CHECK(gcnk_unpad_cols(z.dW2p, W2.grad, H, C, Cp, st));
        GCNK_CHECK(gcnk_matmul_nt(z.dT, Cp, z.W2p, Cp, z.dH1, H, n_loc, Cp, H, st));
        GCNK_CHECK(gcnk_mask_scale_bw(z.dH1, (int64_t)n_loc * H, z.wmask, scale, st));
        gpu_timer_end(TMR_MATMUL_BW);
        gpu_timer_begin(TMR_SPMATMUL_BW);
        GCNK_CHECK(gcnk_matmul_tn(z.AXd, F, z.dH1, H, W1.grad, H, n_loc, F, H, z.mm_ws, z.mm_ws_bytes, st));
        gpu_timer_end(TMR_SPMATMUL_BW);
    }
    finish_pass(training, seq, slot);
}
