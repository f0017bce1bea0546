#include "gcn.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <tuple>

#include "check.h"
#include "gcn_fused.h"
#include "rand.h"
#include "timer.h"

GCNParams GCNParams::get_default() { return {2708, 1433, 16, 7, 0.5f, 0.01f, 5e-4f, 100, 0}; }   // gcn.cpp:9-11

static GCNPlan plan_from_env() {
    const char *s = getenv("GCN_PLAN");
    if (!s || !*s || !strcmp(s, "auto")) return PLAN_AUTO;
    if (!strcmp(s, "modules")) return PLAN_MODULES;
    if (!strcmp(s, "fused")) return PLAN_FUSED;
    fprintf(stderr, "GCN_PLAN must be auto, modules or fused (got '%s')\n", s);
    exit(EXIT_FAILURE);
}

GCN::GCN(GCNParams params_, GCNData *input_data) : params(params_), data(input_data) { build(plan_from_env()); }

GCN::GCN(GCNParams params_, GCNData *input_data, GCNPlan plan, bool quiet) : params(params_), data(input_data), quiet_(quiet) {
    build(plan);
}

GCN::GCN(GCNParams params_, GCNData *input_data, GCNPlan plan, bool quiet, GCNDist dist_)
    : params(params_), data(input_data), dist(dist_), quiet_(quiet) {
    build(plan);
}

void GCN::build(GCNPlan plan) {
    int n_dev = 0;
    GCNK_CHECK(gcnk_device_count(&n_dev));                // no GPU -> fatal: there is no CPU engine behind this class
    init_rand_state();                                    // gcn.cpp:14

    const int N = params.num_nodes, F = params.input_dim, H = params.hidden_dim, C = params.output_dim;
    const size_t nnzX = data->feature_index.indices.size();
    if ((int)data->graph.indptr.size() != N + 1 || (int)data->feature_index.indptr.size() != N + 1 ||
        (int)data->label.size() != N || (int)data->split.size() != N || data->feature_value.size() != nnzX) {
        fprintf(stderr, "GCN: inconsistent input: num_nodes=%d graph rows=%zu feature rows=%zu labels=%zu splits=%zu\n", N,
                data->graph.indptr.size() - 1, data->feature_index.indptr.size() - 1, data->label.size(), data->split.size());
        exit(EXIT_FAILURE);
    }
    for (int i = 0; i < N; i++)
        if (data->split[i] >= 1 && data->split[i] <= 3 && data->label[i] >= 0) split_count[data->split[i]]++;   // truth >= 0 rows (global)

    full_data = data;
    n_loc = N; r0 = 0;
    int symmetric = 0;
    if (dist.world > 1) {
        if (plan == PLAN_MODULES) { fprintf(stderr, "GCN: the row-partitioned engine runs the fused plan only\n"); exit(EXIT_FAILURE); }
        GCNK_CHECK(gcnk_graph_stats(full_data->graph.graph(), nullptr, nullptr, nullptr, &symmetric, nullptr));
        build_partition();                                 // data now points at this rank's row slice
    } else {
        GCNK_CHECK(gcnk_graph_stats(data->graph.graph(), nullptr, nullptr, nullptr, &symmetric, nullptr));
    }

    d_feature_value = upload(data->feature_value);
    d_split = upload(data->split);
    d_label = upload(data->label);
    GCNK_CHECK(gcnk_malloc((void **)&d_truth, sizeof(int) * (size_t)N));

    // Two fused forms: hidden 16-style (layer 2 row-local in one kernel: needs hidden*classes <= 4096) and the WIDE form
    // (gcn_wide.cpp: real GEMMs on the tensor cores, layer 1 re-ordered to (A_hat drop(X)) W1; needs dense features whose
    // width is a multiple of 4 and not above the hidden width).  Both need A_hat = A_hat^T (the backward GraphSum of the
    // reference multiplies by A_hat, not its transpose, module.cpp:103-119) and at most 128 classes.
    const bool narrow_ok = symmetric && C <= 128 && (size_t)H * C <= 4096;
    bool wide_ok = symmetric && C <= 128 && F % 4 == 0 && F <= 1024 && F <= H && nnzX == (size_t)N * F;
    if (wide_ok) {                                         // dense layout: every row stores columns 0..F-1 in order
        const std::vector<int> &fi = full_data->feature_index.indices, &fp = full_data->feature_index.indptr;
        for (int i = 0; i <= N && wide_ok; i++) wide_ok = fp[i] == i * F;
        for (size_t k = 0; k < fi.size() && wide_ok; k += 1 + (k % 97)) wide_ok = fi[k] == (int)(k % (size_t)F);   // sampled: rows are sorted, so a full row of F distinct keys below F is 0..F-1
    }
    const char *fw = getenv("GCN_WIDE");
    const bool force_wide = fw && *fw && strcmp(fw, "0");
    const bool fusable = narrow_ok || wide_ok;
    if (plan == PLAN_AUTO) plan = fusable ? PLAN_FUSED : PLAN_MODULES;
    if (plan == PLAN_FUSED && !fusable) {
        fprintf(stderr, "GCN: the fused plans need a symmetric adjacency and output_dim <= 128, and either hidden*output <= 4096 or dense "
                        "features (width a multiple of 4, at most the hidden width)\n");
        exit(EXIT_FAILURE);
    }
    if (dist.world > 1 && plan != PLAN_FUSED) {
        fprintf(stderr, "GCN: the row-partitioned engine needs a fused plan (symmetric adjacency, output_dim <= 128, and hidden*output "
                        "<= 4096 or dense features no wider than the hidden layer); run this configuration on one GPU\n");
        exit(EXIT_FAILURE);
    }
    plan_ = plan;
    const bool wide = plan == PLAN_FUSED && wide_ok && (!narrow_ok || force_wide);

    modules.reserve(8);
    variables.reserve(8);
    AdamParams adam_params = AdamParams::get_default();
    adam_params.lr = params.learning_rate;
    adam_params.weight_decay = params.weight_decay;

    if (plan_ == PLAN_MODULES) {
        // the reference's network, Variable for Variable and Module for Module (gcn.cpp:20-65)
        variables.emplace_back((int)nnzX, false);
        input = &variables.back();
        modules.push_back(new Dropout(input, params.dropout));
        variables.emplace_back(N * H);
        Variable *layer1_var1 = &variables.back();
        variables.emplace_back(F * H, true);
        Variable *layer1_weight = &variables.back();
        layer1_weight->glorot(F, H);
        modules.push_back(new SparseMatmul(input, layer1_weight, layer1_var1, &data->feature_index, N, F, H));
        variables.emplace_back(N * H);
        Variable *layer1_var2 = &variables.back();
        modules.push_back(new GraphSum(layer1_var1, layer1_var2, &data->graph, H));
        modules.push_back(new ReLU(layer1_var2));
        modules.push_back(new Dropout(layer1_var2, params.dropout));
        variables.emplace_back(N * C);
        Variable *layer2_var1 = &variables.back();
        variables.emplace_back(H * C, true);
        Variable *layer2_weight = &variables.back();
        layer2_weight->glorot(H, C);
        modules.push_back(new Matmul(layer1_var2, layer2_weight, layer2_var1, N, H, C));
        variables.emplace_back(N * C);
        output = &variables.back();
        modules.push_back(new GraphSum(layer2_var1, output, &data->graph, C));
        ce_module = new CrossEntropyLoss(output, d_truth, &loss, C);
        modules.push_back(ce_module);
        optimizer = Adam({{layer1_weight, true}, {layer2_weight, false}}, adam_params);
        return;
    }

    // fused plan: only the weights are Variables; slots 0,1,3,4,6 stay empty so indices match gcn.cpp:21-53
    for (int idx = 0; idx < 7; idx++) {
        if (idx == 2) { variables.emplace_back(F * H, true); variables.back().glorot(F, H); }
        else if (idx == 5) { variables.emplace_back(H * C, true); variables.back().glorot(H, C); }
        else variables.emplace_back(0, false);
    }
    optimizer = Adam({{&variables[2], true}, {&variables[5], false}}, adam_params);

    fz.reset(new Fused);
    fz->wide = wide;
    fz->Cp = (C + 3) / 4 * 4;
    GCNK_CHECK(gcnk_stream_create(&fz->stream));
    const size_t nnzX_loc = data->feature_index.indices.size();
    // gather SOURCES are [N x H] ([N x Cp] in the wide plan; every rank needs all rows: exchanged in place), gather OUTPUTS are local
    const size_t buf = wide ? (size_t)N * fz->Cp : (size_t)N * H;
    const int nbuf = wide ? 2 : 4;
    fz->buf_floats = buf;
    const size_t nh_all = sizeof(float) * buf, nh_loc = sizeof(float) * (size_t)n_loc * H;
    fz->world = dist.world; fz->rank = dist.rank; fz->comm = dist.comm;
    const char *cm = getenv("GCN_COMM"), *ex = getenv("GCN_EXCHANGE");
    fz->signal_exchange = !(ex && !strcmp(ex, "barrier"));
    if (dist.world > 1 && dist.world <= 8 && (wide || H % 4 == 0) && !(cm && !strcmp(cm, "nccl"))) {
        // one slab: [xw_s | h1_s | G | Gm  (wide: T_s | D_s) | flags (128 ints: 0..7 barrier, 64 + 8 b + r buffer b from rank r) |
        // all-reduce area: world slots | loss terms]; export it, import every peer's
        fz->slot_floats = ((size_t)F * H + (size_t)H * C + 4 + 3) / 4 * 4;
        const int max_terms = std::max(split_count[1], std::max(split_count[2], split_count[3]));
        fz->term_region = ((size_t)max_terms + 4 * (size_t)dist.world + 3) / 4 * 4;       // one region per split
        const size_t terms_off = nbuf * buf + 128 + fz->slot_floats * dist.world;         // in floats, a multiple of 4
        const size_t slab_bytes = sizeof(float) * (terms_off + 3 * fz->term_region);
        GCNK_CHECK(gcnk_malloc((void **)&fz->slab, slab_bytes));
        GCNK_CHECK(gcnk_memset(fz->slab, 0, slab_bytes, nullptr));
        GCNK_CHECK(gcnk_malloc((void **)&fz->d_counter, sizeof(unsigned)));
        GCNK_CHECK(gcnk_memset(fz->d_counter, 0, sizeof(unsigned), nullptr));
        GCNK_CHECK(gcnk_stream_sync(nullptr));
        unsigned char mine[64], all[64 * 8];
        GCNK_CHECK(gcnk_ipc_export(fz->slab, mine));
        GCNK_CHECK(gcnk_comm_allgather_bytes(dist.comm, mine, all, 64));
        float failed = 0.f;
        for (int r = 0; r < dist.world; r++) {
            if (r == dist.rank) { fz->peer_slab[r] = fz->slab; continue; }
            if (gcnk_ipc_import(&fz->peer_slab[r], all + 64 * r) != GCNK_OK) { fz->peer_slab[r] = nullptr; failed = 1.f; }
        }
        // all ranks must agree on the transport
        float *d_flag = nullptr;
        GCNK_CHECK(gcnk_malloc((void **)&d_flag, sizeof(float)));
        GCNK_CHECK(gcnk_memcpy_h2d(d_flag, &failed, sizeof(float), nullptr));
        float *bufs[1] = {d_flag};
        const size_t cnt[1] = {1};
        GCNK_CHECK(gcnk_comm_allreduce(dist.comm, bufs, cnt, 1, 1, nullptr));
        GCNK_CHECK(gcnk_memcpy_d2h(&failed, d_flag, sizeof(float), nullptr));
        GCNK_CHECK(gcnk_stream_sync(nullptr));
        GCNK_CHECK(gcnk_free(d_flag));
        fz->p2p = failed == 0.f;
        if (!fz->p2p && !quiet_) fprintf(stderr, "GCN: peer mapping unavailable (%s); using NCCL all-gather\n", gcnk_last_error());
        if (wide) { fz->T_s = fz->slab; fz->D_s = fz->slab + buf; }
        else { fz->xw_s = fz->slab; fz->h1_s = fz->slab + buf; fz->G = fz->slab + 2 * buf; fz->Gm = fz->slab + 3 * buf; }
        for (int r = 0; r < dist.world; r++) {
            fz->flag_arrays[r] = fz->peer_slab[r] ? reinterpret_cast<int *>(static_cast<float *>(fz->peer_slab[r]) + nbuf * buf) : nullptr;
            fz->areas[r] = fz->peer_slab[r] ? static_cast<float *>(fz->peer_slab[r]) + nbuf * buf + 128 : nullptr;
        }
        GCNK_CHECK(gcnk_malloc((void **)&fz->d_err, sizeof(int)));
        GCNK_CHECK(gcnk_memset(fz->d_err, 0, sizeof(int), nullptr));
        GCNK_CHECK(gcnk_malloc_host((void **)&fz->h_err, sizeof(int)));
        *fz->h_err = 0;
        if (fz->p2p) { build_halo(); fz->terms = fz->slab + terms_off; }
    } else if (wide) {
        GCNK_CHECK(gcnk_malloc((void **)&fz->T_s, 2 * nh_all));
        fz->D_s = fz->T_s + buf;
        fz->wide_sources_owned = true;
    } else {
        for (float **p : {&fz->xw_s, &fz->h1_s, &fz->G, &fz->Gm}) GCNK_CHECK(gcnk_malloc((void **)p, nh_all));
    }
    if (!wide) {
        for (float **p : {&fz->P, &fz->dxw}) GCNK_CHECK(gcnk_malloc((void **)p, nh_loc));
        for (int b = 0; b < 2; b++) {
            GCNK_CHECK(gcnk_malloc((void **)&fz->keep0_buf[b], sizeof(uint32_t) * (nnzX_loc / 32 + 4)));
            GCNK_CHECK(gcnk_malloc((void **)&fz->keep1_buf[b], sizeof(uint32_t) * ((size_t)n_loc * H / 32 + 4)));
        }
        fz->keep0 = fz->keep0_buf[0]; fz->keep1 = fz->keep1_buf[0];   // freed through keep0/keep1 (buffer 0) and the [1] entries
        const char *ns = getenv("GCN_NO_RNG_OVERLAP");
        if (!(ns && *ns && strcmp(ns, "0"))) {
            GCNK_CHECK(gcnk_stream_create_low_priority(&fz->rng_stream));   // fills idle slots under the gathers, never ahead of them
            GCNK_CHECK(gcnk_event_create(&fz->ev_ready));
            GCNK_CHECK(gcnk_event_create(&fz->ev_go));
        }
        GCNK_CHECK(gcnk_malloc((void **)&fz->mask, sizeof(uint32_t) * ((size_t)n_loc * gcnk_mask_row_stride_bits(H) / 32 + 4)));
        fz->ws_bytes = gcnk_layer2_workspace(n_loc, H, C);
    } else {
        fz->ws_bytes = gcnk_ce_rows_workspace(n_loc);
    }
    GCNK_CHECK(gcnk_malloc((void **)&fz->ws, fz->ws_bytes));
    GCNK_CHECK(gcnk_malloc((void **)&fz->d_result, sizeof(gcnk_ce_result)));
    GCNK_CHECK(gcnk_malloc((void **)&fz->d_sumsq, sizeof(float)));
    GCNK_CHECK(gcnk_malloc_host((void **)&fz->h_result, sizeof(gcnk_ce_result)));
    GCNK_CHECK(gcnk_malloc_host((void **)&fz->h_sumsq, sizeof(float)));
    GCNK_CHECK(gcnk_malloc_host((void **)&fz->h_red, 8 * sizeof(float)));
    GCNK_CHECK(gcnk_malloc_host((void **)&fz->h_async, sizeof(int)));
    *fz->h_async = 0;
    {
        const char *tl = getenv("GCN_TREE_LOSS");
        fz->seq_loss = !(tl && *tl && strcmp(tl, "0")) && (dist.world == 1 || fz->p2p);
        // measured (r02s/r02t, 1 GPU): right behind layer 2 the one-CTA kernel makes the two backward gathers 40 us slower
        // each; behind the last gather it hides under the weight-gradient kernel.  Row-partitioned, that kernel is too
        // short to hide rank 0's sum of ALL ranks' terms, so there it starts early.
        fz->seq_when = dist.world == 1 ? 1 : 0;
        if (const char *sw = getenv("GCN_SEQ_WHEN")) fz->seq_when = std::max(0, std::min(2, atoi(sw)));
        if (fz->seq_loss) {
            if (!fz->terms) {
                const int max_terms = std::max(split_count[1], std::max(split_count[2], split_count[3]));
                fz->term_region = ((size_t)max_terms + 4 * (size_t)dist.world + 3) / 4 * 4;
                GCNK_CHECK(gcnk_malloc((void **)&fz->terms, sizeof(float) * 3 * fz->term_region));
                GCNK_CHECK(gcnk_memset(fz->terms, 0, sizeof(float) * 3 * fz->term_region, nullptr));
                fz->terms_owned = true;
            }
            GCNK_CHECK(gcnk_malloc((void **)&fz->d_seq, 2 * sizeof(float)));
            GCNK_CHECK(gcnk_malloc_host((void **)&fz->h_seq, 2 * sizeof(float)));
            GCNK_CHECK(gcnk_stream_create(&fz->seq_stream));
            GCNK_CHECK(gcnk_event_create(&fz->ev_l2));
            GCNK_CHECK(gcnk_event_create(&fz->ev_seq));
            for (int sp = 1; sp <= 3; sp++) {
                // Where every local labelled row of split sp stores its loss term: rank r's terms are contiguous (row order)
                // and start at a multiple of 4 floats (so that its push is 16-byte aligned); the gaps hold +0.0f, which a
                // sequential fp32 sum passes over unchanged.  One region per split, so the gaps stay zero.
                std::vector<int> index((size_t)n_loc, 0);
                int off = 0, mine = 0;
                for (int r = 0; r < dist.world; r++) {
                    const int lo = dist.world > 1 ? row_begin[r] : 0, hi = dist.world > 1 ? row_begin[r + 1] : N;
                    int cnt = 0;
                    for (int i = lo; i < hi; i++) cnt += full_data->split[i] == sp && full_data->label[i] >= 0;
                    if (r == dist.rank) { mine = off; fz->term_cnt[sp] = cnt; }
                    off += (cnt + 3) / 4 * 4;
                }
                fz->term_c0[sp] = mine;
                fz->term_len[sp] = off;
                int c = mine;
                for (int i = 0; i < n_loc; i++) {
                    index[i] = c;
                    c += data->split[i] == sp && data->label[i] >= 0;
                }
                fz->term_index[sp] = upload(index);
            }
        }
    }
    GCNK_CHECK(gcnk_rng_create(&fz->slice_rng, 1, 2));
    GCNK_CHECK(gcnk_sum_squares(variables[2].data, variables[2].size, fz->d_sumsq, nullptr));
    GCNK_CHECK(gcnk_memcpy_d2h(fz->h_sumsq, fz->d_sumsq, sizeof(float), nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
    fz->sumsq = *fz->h_sumsq;
    if (wide) { build_wide(); finish_build(); return; }
    gcnk_spmat *sp = data->feature_index.spmat(n_loc, F);     // build the handle (dense detection) up front
    gcnk_graph *g = graph_handle();
    {
        // Exchange fused into the consuming GraphSum (csrc/graph.cu: xgather_kernel): one launch pushes this rank's rows of
        // the source to the peers and aggregates, own columns first.  Needs the rotated row order, set up here before any
        // view of the slice graph exists.  Default up to 4 ranks, where it was measured faster than the separate push kernel +
        // wait at the start of the gather (Reddit shape: +1 % on 2 GPUs, +2 % on 4); on 8 GPUs the own-columns share of a
        // row (1/8) is too small to hide anything and the pushers compete with the gather CTAs: 7 % slower
        // (profiles/r02p_bench_n8*.json).  GCN_FUSED_EXCHANGE=1 / 0 forces it on / off.
        const char *fx = getenv("GCN_FUSED_EXCHANGE");
        const bool want = fx && *fx ? strcmp(fx, "0") != 0 : dist.world <= 4;
        if (dist.world > 1 && fz->p2p && fz->signal_exchange && !fz->use_halo && (H == 16 || H == 12) && want) {
            GCNK_CHECK(gcnk_graph_rotate(g, r0, r0 + n_loc, nullptr));
            fz->fused_xchg = true;
        }
    }

    const char *nv = getenv("GCN_NO_VIEWS"), *na = getenv("GCN_NO_AX");
    fz->use_views = !(nv && *nv && strcmp(nv, "0"));
    if (fz->use_views) {
        for (int s = 1; s <= 3; s++) {
            std::vector<int> keep((size_t)n_loc);
            for (int i = 0; i < n_loc; i++) keep[i] = data->split[i] == s && data->label[i] >= 0;
            fz->keep[s] = upload(keep);
            if (s < 3) GCNK_CHECK(gcnk_graph_create_view(&fz->rows[s], g, fz->keep[s], nullptr, nullptr));   // test split: on first use
        }
        std::vector<int> train_cols((size_t)N);                // column flags are global
        for (int i = 0; i < N; i++) train_cols[i] = full_data->split[i] == 1 && full_data->label[i] >= 0;
        fz->keep[0] = upload(train_cols);
        GCNK_CHECK(gcnk_graph_create_view(&fz->cols_train, g, nullptr, fz->keep[0], nullptr));
    }
    {
        // Opt-in (GCN_OVERLAP=1).  Measured on 2 B200 at Reddit shape (profiles/r02h_*): the exchange leaves the critical path
        // (comm 0.19 -> 0.04 ms per step) but the two extra cross-stream event hops per exchange and the split row sums
        // (every row's chunks are ragged twice) cost more than the hidden push: 570 vs 585 epochs/s.
        const char *ov = getenv("GCN_OVERLAP");
        const bool want = ov && *ov && strcmp(ov, "0");
        if (dist.world > 1 && fz->p2p && fz->signal_exchange && fz->use_views && want) {
            // column blocks of this rank's CSR slice: the columns it owns itself / the columns the peers own, each also
            // restricted to training columns (the backward GraphSum); row-subset views of both per split on top
            std::vector<int> own_f((size_t)N, 0), rem_f((size_t)N, 1), town((size_t)N, 0), trem((size_t)N, 0);
            for (int i = 0; i < N; i++) {
                const bool mine = i >= r0 && i < r0 + n_loc, train = full_data->split[i] == 1 && full_data->label[i] >= 0;
                own_f[i] = mine; rem_f[i] = !mine; town[i] = mine && train; trem[i] = !mine && train;
            }
            int *d_f[4] = {upload(own_f), upload(rem_f), upload(town), upload(trem)};
            GCNK_CHECK(gcnk_graph_create_view(&fz->g_own, g, nullptr, d_f[0], nullptr));
            GCNK_CHECK(gcnk_graph_create_view(&fz->g_rem, g, nullptr, d_f[1], nullptr));
            GCNK_CHECK(gcnk_graph_create_view(&fz->gt_own, g, nullptr, d_f[2], nullptr));
            GCNK_CHECK(gcnk_graph_create_view(&fz->gt_rem, g, nullptr, d_f[3], nullptr));
            for (int s = 1; s <= 2; s++) {
                GCNK_CHECK(gcnk_graph_create_view(&fz->rows_own[s], fz->g_own, fz->keep[s], nullptr, nullptr));
                GCNK_CHECK(gcnk_graph_create_view(&fz->rows_rem[s], fz->g_rem, fz->keep[s], nullptr, nullptr));
            }
            GCNK_CHECK(gcnk_stream_sync(nullptr));
            for (int *f : d_f) GCNK_CHECK(gcnk_free(f));
            GCNK_CHECK(gcnk_malloc((void **)&fz->partial, nh_loc));
            GCNK_CHECK(gcnk_malloc((void **)&fz->d_counter2, sizeof(unsigned)));
            GCNK_CHECK(gcnk_memset(fz->d_counter2, 0, sizeof(unsigned), nullptr));
            GCNK_CHECK(gcnk_stream_create(&fz->comm_stream));
            GCNK_CHECK(gcnk_event_create(&fz->ev_prod));
            fz->overlap = true;
        }
    }
    int dense = 0;
    GCNK_CHECK(gcnk_spmat_is_dense(sp, &dense));
    if (dense && H == 16 && F % 2 == 0 && F <= 1024 && !(na && *na && strcmp(na, "0"))) {
        // AX = A_hat * X for this rank's rows; needs all rows of X once
        float *x_all = d_feature_value;
        if (dist.world > 1) x_all = upload(full_data->feature_value);
        GCNK_CHECK(gcnk_malloc((void **)&fz->AX, sizeof(float) * (size_t)n_loc * F));
        GCNK_CHECK(gcnk_graphsum(g, x_all, fz->AX, F, nullptr));
        GCNK_CHECK(gcnk_stream_sync(nullptr));
        GCNK_CHECK(gcnk_graph_release_scratch(g));            // the [N x F] pre-scaled copy is not needed again
        if (dist.world > 1) GCNK_CHECK(gcnk_free(x_all));
        fz->ax_valid = true;
    }
    const char *nt = getenv("GCN_NO_TMA");
    if (dense && H == 16 && F <= 640 && n_loc > 0 && !(nt && *nt && strcmp(nt, "0"))) {
        // packed (16-byte pitch) copies for the TMA-staged kernels; the unpadded originals stay for inspection
        fz->ld = (F + 31) / 32 * 32;
        GCNK_CHECK(gcnk_malloc((void **)&fz->Xp, sizeof(float) * (size_t)n_loc * fz->ld));
        GCNK_CHECK(gcnk_dense_pack(d_feature_value, n_loc, F, fz->Xp, fz->ld, nullptr));
        if (fz->ax_valid) {
            GCNK_CHECK(gcnk_malloc((void **)&fz->AXp, sizeof(float) * (size_t)n_loc * fz->ld));
            GCNK_CHECK(gcnk_dense_pack(fz->AX, n_loc, F, fz->AXp, fz->ld, nullptr));
            GCNK_CHECK(gcnk_stream_sync(nullptr));
            GCNK_CHECK(gcnk_free(fz->AX));
            fz->AX = nullptr;
        }
        // GCN_TC_TRANSFORM=1: the tcgen05 forms of the two feature-transform products (csrc/matmul_tc.cu) instead of the
        // mma.sync ones (csrc/feature_tma.cu)
        const char *tc = getenv("GCN_TC_TRANSFORM");
        fz->tc_transform = tc && *tc && strcmp(tc, "0") && n_loc >= 2048;
        fz->bw_ws_bytes = std::max(gcnk_dense_transform_bw_workspace(n_loc, F), fz->tc_transform ? gcnk_dense_transform_bw_tc_workspace(n_loc, F, H) : 0);
        GCNK_CHECK(gcnk_malloc((void **)&fz->bw_ws, fz->bw_ws_bytes));
        GCNK_CHECK(gcnk_stream_sync(nullptr));
    }
    finish_build();
}

void GCN::finish_build() {
    // Setup ran on the legacy stream and the engine stream does not synchronise with it: finish everything first.  Then
    // meet the other ranks, so that nobody enters its first exchange while another rank is still building (uneven
    // build times would otherwise eat into the flag-wait timeout).
    GCNK_CHECK(gcnk_device_sync());
    if (dist.world > 1) {
        float *b[1] = {fz->ws};
        const size_t c1[1] = {1};
        GCNK_CHECK(gcnk_comm_allreduce(dist.comm, b, c1, 1, 1, nullptr));
        GCNK_CHECK(gcnk_device_sync());
    }
}

// Halo exchange: peer p only ever reads the rows of a gather source that its columns reference.  Every rank
// marks the columns of its own CSR slice in an N-bit map, the maps are all-gathered once, and each rank keeps, per
// peer, the list of its own rows that peer needs.  The lists replace the contiguous push when they save at least 10 %
// of the rows for some peer (GCN_HALO=1 forces them, =0 disables them); on the dense synthetic graphs every peer
// references nearly every row and the contiguous (fully coalesced) push stays.
void GCN::build_halo() {
    Fused &z = *fz;
    const int N = params.num_nodes;
    const size_t bytes = ((size_t)N + 7) / 8;
    std::vector<unsigned char> mine(bytes, 0), all(bytes * (size_t)dist.world, 0);
    for (int c : data->graph.indices) mine[(size_t)c >> 3] |= (unsigned char)(1u << (c & 7));
    GCNK_CHECK(gcnk_comm_allgather_bytes(dist.comm, mine.data(), all.data(), (int)bytes));
    const char *hv = getenv("GCN_HALO");
    bool any_saving = false;
    std::vector<std::vector<int>> lists((size_t)dist.world);
    for (int p = 0; p < dist.world; p++) {
        if (p == dist.rank) continue;
        const unsigned char *map = all.data() + bytes * (size_t)p;
        for (int i = r0; i < r0 + n_loc; i++)
            if ((map[(size_t)i >> 3] >> (i & 7)) & 1u) lists[p].push_back(i - r0);
        if ((double)lists[p].size() < 0.9 * (double)n_loc) any_saving = true;
    }
    z.use_halo = hv && *hv ? atoi(hv) != 0 : any_saving;
    if (!z.use_halo) return;
    for (int p = 0; p < dist.world; p++) {
        if (p == dist.rank) continue;
        z.halo_count[p] = (int)lists[p].size();
        z.halo_rows[p] = upload(lists[p]);
    }
}

gcnk_stream_t GCN::engine_stream() const { return fz ? fz->stream : nullptr; }

gcnk_graph *GCN::graph_handle() {
    return dist.world > 1 ? data->graph.graph_slice(params.num_nodes, d_dinv_global) : data->graph.graph();
}

void slice_rows(const GCNData &src, int r0, int n_rows, GCNData &dst) {
    auto slice_csr = [&](const SparseIndex &a, SparseIndex &b) {
        const int e0 = a.indptr[r0], e1 = a.indptr[r0 + n_rows];
        b.indptr.resize((size_t)n_rows + 1);
        for (int i = 0; i <= n_rows; i++) b.indptr[i] = a.indptr[r0 + i] - e0;
        b.indices.assign(a.indices.begin() + e0, a.indices.begin() + e1);
    };
    slice_csr(src.graph, dst.graph);
    slice_csr(src.feature_index, dst.feature_index);
    const int x0 = src.feature_index.indptr[r0], x1 = src.feature_index.indptr[r0 + n_rows];
    dst.feature_value.assign(src.feature_value.begin() + x0, src.feature_value.begin() + x1);
    dst.split.assign(src.split.begin() + r0, src.split.begin() + r0 + n_rows);
    dst.label.assign(src.label.begin() + r0, src.label.begin() + r0 + n_rows);
}

// Row partition: this rank keeps rows [r0, r0 + n_loc) of the graph (column ids stay global), of X, labels and split.
void GCN::build_partition() {
    const int N = params.num_nodes;
    gcnk_graph *gfull = full_data->graph.graph();
    const float *dinv_full = nullptr;
    GCNK_CHECK(gcnk_graph_dinv(gfull, &dinv_full));
    GCNK_CHECK(gcnk_malloc((void **)&d_dinv_global, sizeof(float) * (size_t)N));
    GCNK_CHECK(gcnk_memcpy_d2d(d_dinv_global, dinv_full, sizeof(float) * (size_t)N, nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
    full_data->graph.release_device();                      // the full CSR is not needed on the device any more

    row_begin.assign((size_t)dist.world + 1, 0);
    GCNK_CHECK(gcnk_partition_rows(full_data->graph.indptr.data(), N, dist.world, row_begin.data()));
    r0 = row_begin[dist.rank];
    n_loc = row_begin[dist.rank + 1] - r0;

    local.reset(new GCNData);
    slice_rows(*full_data, r0, n_loc, *local);
    data = local.get();
}

GCN::~GCN() {
    if (copy_stream) { gcnk_stream_sync(copy_stream); gcnk_stream_destroy(copy_stream); }   // an upload may still be in flight
    if (ev_copied) gcnk_event_destroy(ev_copied);
    for (auto m : modules) delete m;
    fz.reset();                                             // views before the graph they borrow from
    for (void *p : {(void *)d_truth, (void *)d_split, (void *)d_label, (void *)d_feature_value, (void *)d_feature_spare,
                    (void *)d_dinv_global})
        if (p) gcnk_free(p);
}

void GCN::set_input_from_host(const float *h_values) {
    consume_pending_input();                                // a prefetched input is older than this one: retire it first
    const size_t bytes = sizeof(float) * data->feature_index.indices.size();
    if (fz && fz->wide && dist.world > 1) {
        // The row-partitioned wide plan reads every node's features (gcn_wide.cpp): this rank's rows go into its slice of the
        // replicated matrix and the slices are all-gathered over NVLink (NCCL; 980 MB at products shape).
        const int F = params.input_dim;
        GCNK_CHECK(gcnk_memcpy_h2d(fz->X_all + (size_t)r0 * F, h_values, bytes, engine_stream()));
        GCNK_CHECK(gcnk_comm_allgather_rows(dist.comm, fz->X_all, row_begin.data(), F, engine_stream()));
        fz->ax_valid = false;
        return;
    }
    GCNK_CHECK(gcnk_memcpy_h2d(d_feature_value, h_values, bytes, engine_stream()));
    if (fz) { fz->ax_valid = false; fz->xp_dirty = fz->Xp != nullptr; }   // A_hat*X is stale (eval falls back to the gather path); re-pack X
}

// Upload into the spare buffer on the copy stream.  The spare buffer is not read by anything in flight: it was the
// input of a pass that has completed (every pass ends with a host synchronisation before its results are returned).
void GCN::start_input_upload(const float *h_values) {
    const size_t bytes = sizeof(float) * data->feature_index.indices.size();
    const bool replicated = fz && fz->wide && dist.world > 1;             // see set_input_from_host
    if (!copy_stream) {
        GCNK_CHECK(gcnk_stream_create(&copy_stream));
        GCNK_CHECK(gcnk_event_create(&ev_copied));
        if (replicated) GCNK_CHECK(gcnk_malloc((void **)&fz->X_all_spare, sizeof(float) * (size_t)params.num_nodes * params.input_dim));
        else GCNK_CHECK(gcnk_malloc((void **)&d_feature_spare, std::max<size_t>(bytes, sizeof(float))));
    }
    if (input_pending) GCNK_CHECK(gcnk_stream_sync(copy_stream));   // never two uploads into the one spare buffer
    if (replicated) {
        const int F = params.input_dim;
        GCNK_CHECK(gcnk_memcpy_h2d(fz->X_all_spare + (size_t)r0 * F, h_values, bytes, copy_stream));
        // the slices meet on the copy stream too, under the passes in flight — unless the passes themselves use NCCL (the
        // fallback transport): two streams must not drive one communicator at a time, so the gather then waits for consume
        fz->spare_needs_gather = !fz->p2p;
        if (fz->p2p) GCNK_CHECK(gcnk_comm_allgather_rows(dist.comm, fz->X_all_spare, row_begin.data(), F, copy_stream));
    } else {
        GCNK_CHECK(gcnk_memcpy_h2d(d_feature_spare, h_values, bytes, copy_stream));
    }
    GCNK_CHECK(gcnk_event_record(ev_copied, copy_stream));
    input_pending = true;
}

// Called at the start of every pass: switch to the prefetched input once its upload has finished (device-side wait).
void GCN::consume_pending_input() {
    if (!input_pending) return;
    GCNK_CHECK(gcnk_stream_wait_event(engine_stream(), ev_copied));
    input_pending = false;
    if (fz && fz->wide && dist.world > 1) {
        std::swap(fz->X_all, fz->X_all_spare);
        if (fz->spare_needs_gather) GCNK_CHECK(gcnk_comm_allgather_rows(dist.comm, fz->X_all, row_begin.data(), params.input_dim, engine_stream()));
        fz->spare_needs_gather = false;
        fz->ax_valid = false;
        return;
    }
    std::swap(d_feature_value, d_feature_spare);
    if (fz) { fz->ax_valid = false; fz->xp_dirty = fz->Xp != nullptr; }
    if (fz && fz->wide && !fz->x_all_owned) fz->X_all = d_feature_value;   // single-GPU wide plan: X_all aliases the feature buffer
}

void GCN::epoch_prefetch(int eval_split, const float *h_next, float *train_loss, float *train_acc, float *eval_loss, float *eval_acc) {
    if (plan_ != PLAN_FUSED) {                              // the modules plan synchronises after every operator: no overlap to win
        std::tie(*train_loss, *train_acc) = train_epoch();
        std::tie(*eval_loss, *eval_acc) = eval(eval_split);
        start_input_upload(h_next);
        return;
    }
    fused_enqueue(1, true, 0);
    fused_enqueue(eval_split, false, 1);
    start_input_upload(h_next);                             // runs under the two passes just enqueued
    std::tie(*train_loss, *train_acc) = fused_collect(0, true);
    train_count = last_count; train_wrong = last_wrong;
    std::tie(*eval_loss, *eval_acc) = fused_collect(1, false);
}

// ---------------------------------------------------------------------------- modules plan ----
void GCN::set_input() {
    // restores the feature values the in-place Dropout of the previous pass destroyed (gcn.cpp:73-76);
    // device-to-device here, where the reference GPU path re-uploads them from the host (cuda_gcn.cu:81-83)
    consume_pending_input();
    GCNK_CHECK(gcnk_memcpy_d2d(input->data, d_feature_value, sizeof(float) * (size_t)input->size, nullptr));
}

void GCN::set_truth(int current_split) {
    GCNK_CHECK(gcnk_set_truth(d_truth, d_split, d_label, current_split, params.num_nodes, nullptr));
}

float GCN::get_accuracy() {
    // counted on the device inside the CrossEntropyLoss forward with the reference's rule (gcn.cpp:83-96)
    last_count = ce_module->last_count;
    last_wrong = ce_module->last_wrong;
    return float(last_count - last_wrong) / last_count;
}

float GCN::get_l2_penalty() {
    static thread_local float *d_out = nullptr;
    if (!d_out) GCNK_CHECK(gcnk_malloc((void **)&d_out, sizeof(float)));
    float l2 = 0;
    GCNK_CHECK(gcnk_sum_squares(variables[2].data, variables[2].size, d_out, nullptr));
    GCNK_CHECK(gcnk_memcpy_d2h(&l2, d_out, sizeof(float), nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
    return params.weight_decay * l2 / 2;
}

// ------------------------------------------------------------------------------ fused plan ----
// Before a producer of gather source `d_all`: register the peers' copies so its epilogue mirrors the rows
// (opt-in, GCNK_MIRROR_EPILOGUE=1; by default the registration stays pending and publish() pushes the rows).
void GCN::mirror(float *d_all, int dim) {
    if (dist.world <= 1 || !fz->p2p) return;
    Fused &z = *fz;
    float *peers[8];
    int n = 0;
    const size_t off = (size_t)(d_all - z.slab) + (size_t)r0 * dim;
    for (int r = 0; r < dist.world; r++)
        if (r != dist.rank) peers[n++] = static_cast<float *>(z.peer_slab[r]) + off;
    GCNK_CHECK(gcnk_mirror_next(d_all + (size_t)r0 * dim, peers, n));
}

// After the producer of gather source `d_all` (buffer b of the slab): make this rank's rows available to every rank.
//   peer memory, default: one push kernel copies the finished rows (all of them, or per peer only the rows that peer's
//     columns reference — halo exchange) into the peers' slabs over NVLink and publishes seq[b] in their flag slots;
//     nothing waits here.  await(b) then arms the consuming gather, which checks the flags at its own start.
//   GCN_EXCHANGE=barrier: the round-1 form, push + flag barrier in one launch (every rank waits for every rank here).
//   GCN_COMM=nccl / no peer mapping: NCCL all-gather (grouped broadcasts).
void GCN::publish(float *d_all, int dim, bool on_comm_stream) {
    if (dist.world <= 1) return;
    Fused &z = *fz;
    gcnk_stream_t ps = on_comm_stream ? z.comm_stream : z.stream;
    unsigned *counter = on_comm_stream ? z.d_counter2 : z.d_counter;   // the push kernels' last-CTA counter: one per stream
    if (!on_comm_stream) gpu_timer_begin(TMR_COMM);
    if (z.p2p) {
        const int b = (int)((size_t)(d_all - z.slab) / z.buf_floats);
        float *own = d_all + (size_t)r0 * dim;
        // hidden-16 plan: a producer with a mirrored epilogue (opt-in) has consumed the registration made by mirror() and
        // stored the rows remotely itself; the wide plan's producers never do
        const bool mirrored = !z.wide && !gcnk_mirror_pending(own);
        float *peers[8];
        int *slots[8];
        const int *lists[8];
        int counts[8], n = 0;
        const size_t off = (size_t)(d_all - z.slab) + (size_t)r0 * dim;
        for (int r = 0; r < dist.world; r++) {
            if (r == dist.rank) continue;
            peers[n] = static_cast<float *>(z.peer_slab[r]) + off;
            slots[n] = z.flag_arrays[r] + 64 + 8 * b + dist.rank;
            lists[n] = z.halo_now ? z.halo_now->rows[r] : z.halo_rows[r];
            counts[n] = z.halo_now ? z.halo_now->count[r] : z.halo_count[r];
            n++;
        }
        const bool listed = !mirrored && (z.halo_now ? z.halo_now->valid : z.use_halo);
        z.halo_now = nullptr;
        if (z.signal_exchange) {
            ++z.seq[b];
            GCNK_CHECK(gcnk_peer_push_signal(own, peers, n, mirrored ? 0 : (size_t)n_loc * dim, listed ? lists : nullptr, counts, dim,
                                             slots, z.seq[b], counter, ps));
        } else if (!mirrored) {
            GCNK_CHECK(gcnk_peer_push_barrier(own, peers, n, (size_t)n_loc * dim, z.flag_arrays, dist.rank, dist.world, ++z.barrier_value,
                                              z.d_err, z.d_counter, z.stream));
        } else {
            GCNK_CHECK(gcnk_peer_barrier(z.flag_arrays, dist.rank, dist.world, ++z.barrier_value, z.d_err, z.stream));
        }
    } else {
        GCNK_CHECK(gcnk_comm_allgather_rows(dist.comm, d_all, row_begin.data(), dim, z.stream));
    }
    if (!on_comm_stream) gpu_timer_end(TMR_COMM);
}

// Exchange of gather source `buf` overlapped with the local part of the GraphSum that consumes it (north star:
// "communication overlapped with local-row aggregation"): the push of this rank's rows to the peers runs on the
// communication stream while the compute stream already aggregates the columns this rank owns (view v_own; raw partial row
// sums).  The rest of the sum — the columns other ranks own (v_rem) — starts from those partials and waits, inside the
// kernel, for the peers' arrival flags.  Own columns first, remote second: a fixed order.  `final_gather(view)` launches
// the consumer (with its epilogue) on the given view.  Returns false when the overlap is off: the caller runs the plain
// publish / await / gather sequence.
bool GCN::exchange_overlapped(float *buf, int dim, gcnk_graph *v_own, gcnk_graph *v_rem) {
    Fused &z = *fz;
    if (!z.overlap || !v_own || !v_rem) return false;
    GCNK_CHECK(gcnk_event_record(z.ev_prod, z.stream));
    GCNK_CHECK(gcnk_stream_wait_event(z.comm_stream, z.ev_prod));
    publish(buf, dim, true);
    GCNK_CHECK(gcnk_gather_raw(v_own, buf, z.partial, dim, z.stream));
    await(buf, dim);
    GCNK_CHECK(gcnk_gather_init_next(z.partial));
    return true;
}

// Fused form of publish + await: arms the next gather launch to push this rank's rows of `d_all` itself and to wait for the
// peers' rows as it reaches their columns (gcnk_gather_exchange_next).  false: not in use, take the two-step path.
bool GCN::arm_exchange(float *d_all, int dim) {
    Fused &z = *fz;
    if (dist.world <= 1 || !z.fused_xchg) return false;
    const int b = (int)((size_t)(d_all - z.slab) / z.buf_floats);
    float *peers[8];
    int *slots[8], n = 0;
    const size_t off = (size_t)(d_all - z.slab) + (size_t)r0 * dim;
    for (int r = 0; r < dist.world; r++) {
        if (r == dist.rank) continue;
        peers[n] = static_cast<float *>(z.peer_slab[r]) + off;
        slots[n] = z.flag_arrays[r] + 64 + 8 * b + dist.rank;
        n++;
    }
    ++z.seq[b];
    GCNK_CHECK(gcnk_gather_exchange_next(d_all + (size_t)r0 * dim, (size_t)n_loc * dim, peers, n, slots, z.flag_arrays[dist.rank] + 64 + 8 * b,
                                         dist.rank, dist.world, z.seq[b], z.d_counter, z.d_err));
    return true;
}

// Arms the next gather launch: it reads buffer `d_all`, so it must see every rank's rows of production seq[b].
void GCN::await(float *d_all, int dim) {
    Fused &z = *fz;
    if (dist.world <= 1 || !z.p2p || !z.signal_exchange) return;
    const int b = (int)((size_t)(d_all - z.slab) / z.buf_floats);
    (void)dim;
    GCNK_CHECK(gcnk_gather_wait_next(z.flag_arrays[dist.rank] + 64 + 8 * b, dist.world, dist.rank, z.seq[b], z.d_err));
}

// The reference's own summation order for the printed loss (module.cpp:125-143): the per-row loss terms of split `sidx`
// (written by the layer-2 / CE kernel just enqueued, into the split's region of the terms buffer) are added up by
// gcnk_sequential_sum into ws[3] — the value the pass reports (ws[0] keeps the parallel sum).
// Row-partitioned: only rank 0 adds (the other ranks push their range of terms to rank 0 — one 16-byte-aligned copy to
// ONE peer — and contribute 0 to the sum of ws[0] over ranks that ends every pass, so every rank still reports the same
// number, bit-identical to a single-GPU run).  Training: on a side stream, under the backward pass.
void GCN::enqueue_loss_sum(int sidx_l, bool training, int slot) {
    Fused &z = *fz;
    gcnk_stream_t st = z.stream;
    (void)slot;
    float *region = z.terms + (size_t)(sidx_l - 1) * z.term_region;
    if (dist.world > 1 && dist.rank != 0) {
        float *peer[1] = {static_cast<float *>(z.peer_slab[0]) + (size_t)(region - z.slab) + z.term_c0[sidx_l]};
        int *slot_flag[1] = {z.flag_arrays[0] + 64 + 8 * 4 + dist.rank};
        ++z.seq[4];
        GCNK_CHECK(gcnk_peer_push_signal(region + z.term_c0[sidx_l], peer, 1, (size_t)(z.term_cnt[sidx_l] + 3) / 4 * 4, nullptr, nullptr, 1,
                                         slot_flag, z.seq[4], z.d_counter, st));
        return;                                                          // ws[3] stays 0 here: rank 0 supplies the sum
    }
    if (training && z.seq_when != 0) { z.seq_pending = sidx_l; return; }   // launched later in the pass: flush_loss_sum
    launch_loss_sum(sidx_l, training);
}

void GCN::flush_loss_sum() {
    Fused &z = *fz;
    if (!z.seq_pending) return;
    launch_loss_sum(z.seq_pending, true);
    z.seq_pending = 0;
}

void GCN::launch_loss_sum(int sidx_l, bool training) {
    Fused &z = *fz;
    gcnk_stream_t st = z.stream;
    float *region = z.terms + (size_t)(sidx_l - 1) * z.term_region;
    const int *flags = nullptr;
    if (dist.world > 1) { ++z.seq[4]; flags = z.flag_arrays[0] + 64 + 8 * 4; }
    const bool side = training && z.seq_when != 2;
    gcnk_stream_t ss = side ? z.seq_stream : st;                        // eval: nothing follows that could hide it
    if (side) {
        GCNK_CHECK(gcnk_event_record(z.ev_l2, st));
        GCNK_CHECK(gcnk_stream_wait_event(z.seq_stream, z.ev_l2));
    }
    // into ws[3], the slot the parallel reduction leaves at 0 on every rank: the sum over ranks that ends the pass carries
    // it to everybody unchanged (S + 0 + ... + 0)
    GCNK_CHECK(gcnk_sequential_sum(region, z.term_len[sidx_l], z.ws + 3, 0.f, flags, flags ? dist.world : 0, 0, z.seq[4], z.d_err, ss));
    if (side) {
        GCNK_CHECK(gcnk_event_record(z.ev_seq, z.seq_stream));
        z.seq_joined = false;
    }
}

// The end of every fused pass: cross-rank sums, the scalars to the host, the optimiser step.
void GCN::finish_pass(bool training, bool seq, int slot) {
    Fused &z = *fz;
    gcnk_stream_t st = z.stream;
    Variable &W1 = variables[2], &W2 = variables[5];
    // the pass is complete only with its loss; and (row-partitioned) no peer may overwrite the loss terms in this rank's
    // slab — which it can do as soon as it has passed the barrier below — before they have been added up
    flush_loss_sum();
    if (!z.seq_joined) { GCNK_CHECK(gcnk_stream_wait_event(st, z.ev_seq)); z.seq_joined = true; }
    if (dist.world > 1) {
        // sums over nodes: dW1, dW2 and {sum of loss terms, count, wrong}; every rank then applies the same update.
        // This is also the one true barrier of the pass: nobody starts the next pass (and overwrites a gather source
        // in a peer's slab) before every rank has finished reading this pass's sources.
        float *bufs[3] = {z.ws, W1.grad, W2.grad};
        const size_t counts[3] = {4, (size_t)W1.size, (size_t)W2.size};
        gpu_timer_begin(TMR_COMM);
        if (z.p2p)
            GCNK_CHECK(gcnk_peer_allreduce(bufs, counts, training ? 3 : 1, z.areas, z.slot_floats, z.flag_arrays, dist.rank, dist.world,
                                           ++z.barrier_value, z.d_err, z.d_counter, st));
        else
            GCNK_CHECK(gcnk_comm_allreduce(dist.comm, bufs, counts, training ? 3 : 1, 0, st));
        gpu_timer_end(TMR_COMM);
    }
    GCNK_CHECK(gcnk_memcpy_d2h(z.h_red + 4 * slot, z.ws, 4 * sizeof(float), st));
    z.sumsq_used[slot] = -1.f;                                            // eval: the penalty of the weights as they are at collect time
    if (training) {
        z.sumsq_used[slot] = z.sumsq;                                     // gcn.cpp:98-105: W1 as it was in this forward
        optimizer.step(z.d_sumsq, st);
        GCNK_CHECK(gcnk_memcpy_d2h(z.h_sumsq, z.d_sumsq, sizeof(float), st));
        z.sumsq_pending = true;
    }
}

// Enqueues one pass on the stream; nothing is read back until fused_collect.  slot: which pinned result slot to use.
void GCN::fused_enqueue(int current_split, bool training, int slot) {
    if (fz->wide) { wide_enqueue(current_split, training, slot); return; }
    consume_pending_input();
    Fused &z = *fz;
    gcnk_stream_t st = z.stream;
    gpu_timer_set_stream(st);
    const int N = params.num_nodes, F = params.input_dim, H = params.hidden_dim, C = params.output_dim;
    const int64_t nnzX_loc = (int64_t)data->feature_index.indices.size();
    const int64_t nnzX_all = (int64_t)full_data->feature_index.indices.size();
    const int64_t x_off = full_data->feature_index.indptr[r0];           // this rank's first feature entry in the global order
    const float p = params.dropout;
    const float scale = 1 / (1 - p);                                      // module.cpp:212
    const bool drop = training && (int)(p * (float)MY_RAND_MAX) > 0;      // threshold 0 keeps everything
    gcnk_graph *g = graph_handle();
    gcnk_spmat *sp = data->feature_index.spmat(n_loc, F);
    const float *dinv = nullptr;                                          // d^-1/2 of the local rows
    GCNK_CHECK(gcnk_graph_dinv(g, &dinv));
    Variable &W1 = variables[2], &W2 = variables[5];
    const size_t own = (size_t)r0 * H;                                    // this rank's slice of an [N x H] gather source

    gcnk_graph *g_rows = g, *g_cols = g;
    if (z.use_views) {
        const int sidx = current_split >= 1 && current_split <= 3 ? current_split : 0;
        if (sidx && !z.rows[sidx]) {
            GCNK_CHECK(gcnk_graph_create_view(&z.rows[sidx], g, z.keep[sidx], nullptr, nullptr));
            GCNK_CHECK(gcnk_stream_sync(nullptr));
        }
        if (sidx) g_rows = z.rows[sidx];
        g_cols = z.cols_train;
    }
    auto gather_timer = [&](gcnk_graph *view) { return view == g ? TMR_GATHER_FULL : TMR_GATHER_PART; };

    // The reference draws nnz(X) then N*H values per training pass from ONE stream in element order
    // (module.cpp:214-218 via gcn.cpp:110-111).  Each rank jumps a copy of the stream to its own rows, so the
    // masks are the same bits whatever the partition; the shared stream then advances by the global counts.
    auto draw_masks = [&](const uint64_t *state, uint32_t *k0, uint32_t *k1, gcnk_stream_t stream) {
        GCNK_CHECK(gcnk_rng_set_state(z.slice_rng, state[0], state[1]));
        GCNK_CHECK(gcnk_rng_skip(z.slice_rng, (uint64_t)x_off));
        GCNK_CHECK(gcnk_dropout_mask(z.slice_rng, k0, nnzX_loc, p, stream));
        GCNK_CHECK(gcnk_rng_set_state(z.slice_rng, state[0], state[1]));
        GCNK_CHECK(gcnk_rng_skip(z.slice_rng, (uint64_t)nnzX_all + (uint64_t)r0 * H));
        GCNK_CHECK(gcnk_dropout_mask(z.slice_rng, k1, (int64_t)n_loc * H, p, stream));
    };
    if (training) {
        gpu_timer_begin(TMR_DROPOUT_FW);
        uint64_t state[2];
        GCNK_CHECK(gcnk_rng_get_state(global_rng(), state));
        if (drop) {
            z.keep0 = z.keep0_buf[z.cur]; z.keep1 = z.keep1_buf[z.cur];
            const bool ahead = z.pre_valid && z.pre_state[0] == state[0] && z.pre_state[1] == state[1];
            // Either way the side stream's last draw targets THIS buffer pair: wait for it before reading the bits — or
            // before redrawing them in line, so that a stale draw still in flight cannot overwrite the fresh bits.
            if (z.pre_valid) GCNK_CHECK(gcnk_stream_wait_event(st, z.ev_ready));
            if (!ahead) draw_masks(state, z.keep0, z.keep1, st);
            z.pre_valid = false;
        }
        GCNK_CHECK(gcnk_rng_skip(global_rng(), (uint64_t)nnzX_all + (uint64_t)N * H));   // consumed even when p == 0
        gpu_timer_end(TMR_DROPOUT_FW);
    }

    if (!training && z.ax_valid && (z.AX || z.AXp)) {
        // eval: A_hat*(X*W1) = (A_hat*X)*W1, ReLU and the pre-scale for the next gather in the epilogue
        gpu_timer_begin(TMR_SPMATMUL_FW);
        mirror(z.h1_s, H);
        if (z.AXp && z.tc_transform) GCNK_CHECK(gcnk_dense_transform_tc(z.AXp, z.ld, n_loc, F, W1.data, z.h1_s + own, H, nullptr, 1.0f, dinv, 1, st));
        else if (z.AXp) GCNK_CHECK(gcnk_dense_transform_ld(z.AXp, z.ld, n_loc, F, W1.data, z.h1_s + own, H, nullptr, 1.0f, dinv, 1, st));
        else GCNK_CHECK(gcnk_dense_transform(z.AX, n_loc, F, W1.data, z.h1_s + own, H, nullptr, 1.0f, dinv, 1, st));
        gpu_timer_end(TMR_SPMATMUL_FW);
    } else {
        // M0 Dropout + M1 SparseMatmul: keep bits applied on read; the stored feature values are never modified,
        // so no set_input() copy is needed
        gpu_timer_begin(TMR_SPMATMUL_FW);
        if (z.xp_dirty) { GCNK_CHECK(gcnk_dense_pack(d_feature_value, n_loc, F, z.Xp, z.ld, st)); z.xp_dirty = false; }
        mirror(z.xw_s, H);
        if (z.Xp && z.tc_transform) GCNK_CHECK(gcnk_dense_transform_tc(z.Xp, z.ld, n_loc, F, W1.data, z.xw_s + own, H, drop ? z.keep0 : nullptr, scale, dinv, 0, st));
        else if (z.Xp) GCNK_CHECK(gcnk_dense_transform_ld(z.Xp, z.ld, n_loc, F, W1.data, z.xw_s + own, H, drop ? z.keep0 : nullptr, scale, dinv, 0, st));
        else GCNK_CHECK(gcnk_spmm_fw(sp, d_feature_value, W1.data, z.xw_s + own, H, drop ? z.keep0 : nullptr, scale, dinv, st));
        gpu_timer_end(TMR_SPMATMUL_FW);
        if (drop && z.rng_stream) {
            // Draw the NEXT training pass's masks now, on the side stream, into the other buffer (its last readers
            // were in the previous training pass, which has completed: every pass ends with a host sync).  It is
            // released only after the feature transform above (issue-bound, like the generator) so that it runs
            // under the L2-bound gathers.
            GCNK_CHECK(gcnk_event_record(z.ev_go, st));
            GCNK_CHECK(gcnk_stream_wait_event(z.rng_stream, z.ev_go));
            GCNK_CHECK(gcnk_rng_get_state(global_rng(), z.pre_state));
            draw_masks(z.pre_state, z.keep0_buf[z.cur ^ 1], z.keep1_buf[z.cur ^ 1], z.rng_stream);
            GCNK_CHECK(gcnk_event_record(z.ev_ready, z.rng_stream));
            z.pre_valid = true;
            z.cur ^= 1;
        }
        // M2 GraphSum + M3 ReLU + M4 Dropout in the gather's epilogue
        gpu_timer_begin(TMR_GATHER_FULL);
        gcnk_graph *v = g;
        if (arm_exchange(z.xw_s, H)) {}
        else if (exchange_overlapped(z.xw_s, H, z.g_own, z.g_rem)) v = z.g_rem;
        else { publish(z.xw_s, H); await(z.xw_s, H); }
        mirror(z.h1_s, H);
        GCNK_CHECK(gcnk_gather_relu_drop(v, z.xw_s, z.h1_s + own, drop ? z.keep1 : nullptr, training ? z.mask : nullptr,
                                         training ? scale : 1.0f, H, st));
        gpu_timer_end(TMR_GATHER_FULL);
    }
    // the layer-2 aggregation at width H, only for the rows whose logits the loss looks at
    gpu_timer_begin(gather_timer(g_rows));
    {
        const int sv = current_split >= 1 && current_split <= 3 ? current_split : 0;
        gcnk_graph *v = g_rows;
        if (sv && z.overlap && !z.rows_own[sv]) {              // test split: on first use
            GCNK_CHECK(gcnk_graph_create_view(&z.rows_own[sv], z.g_own, z.keep[sv], nullptr, nullptr));
            GCNK_CHECK(gcnk_graph_create_view(&z.rows_rem[sv], z.g_rem, z.keep[sv], nullptr, nullptr));
            GCNK_CHECK(gcnk_stream_sync(nullptr));
        }
        if (arm_exchange(z.h1_s, H)) {}
        else if (exchange_overlapped(z.h1_s, H, sv ? z.rows_own[sv] : z.g_own, sv ? z.rows_rem[sv] : z.g_rem)) v = sv ? z.rows_rem[sv] : z.g_rem;
        else { publish(z.h1_s, H); await(z.h1_s, H); }
        GCNK_CHECK(gcnk_gather_plain(v, z.h1_s, z.P, H, st));
    }
    gpu_timer_end(gather_timer(g_rows));

    // M5 Matmul + M7 CrossEntropyLoss + get_accuracy (+ Matmul backward when training), row-local.
    // count = labelled rows of the split over ALL ranks: the gradient is divided by it (module.cpp:154-158)
    gpu_timer_begin(TMR_LOSS_FW);
    if (training) mirror(z.G, H);
    const int sidx_l = current_split >= 1 && current_split <= 3 ? current_split : 0;
    const bool seq = z.seq_loss && sidx_l != 0 && split_count[sidx_l] >= SEQ_LOSS_MIN_ROWS;
    GCNK_CHECK(gcnk_layer2_fused_terms(z.P, W2.data, d_split, d_label, current_split, n_loc, H, C, training,
                                       split_count[current_split & 3], dinv, training ? z.G + own : nullptr,
                                       training ? W2.grad : nullptr, nullptr, z.d_result, z.ws, z.ws_bytes,
                                       seq ? z.terms + (size_t)(sidx_l - 1) * z.term_region : nullptr, seq ? z.term_index[sidx_l] : nullptr, st));
    gpu_timer_end(TMR_LOSS_FW);
    if (seq) enqueue_loss_sum(sidx_l, training, slot);
    z.seq_used[slot] = seq;

    if (training) {
        // backward of M6/M5 is inside layer2; M4/M3/M2 backward = one masked gather + one plain gather
        gpu_timer_begin(gather_timer(g_cols));
        gcnk_graph *v = g_cols;
        if (arm_exchange(z.G, H)) {}
        else if (exchange_overlapped(z.G, H, z.gt_own, z.gt_rem)) v = z.gt_rem;
        else { publish(z.G, H); await(z.G, H); }
        mirror(z.Gm, H);
        GCNK_CHECK(gcnk_gather_mask(v, z.G, z.Gm + own, z.mask, scale, H, st));
        gpu_timer_end(gather_timer(g_cols));
        gpu_timer_begin(TMR_GATHER_FULL);
        v = g;
        if (arm_exchange(z.Gm, H)) {}
        else if (exchange_overlapped(z.Gm, H, z.g_own, z.g_rem)) v = z.g_rem;
        else { publish(z.Gm, H); await(z.Gm, H); }
        GCNK_CHECK(gcnk_gather_plain(v, z.Gm, z.dxw, H, st));
        gpu_timer_end(TMR_GATHER_FULL);
        if (z.seq_when == 1) flush_loss_sum();                           // under the weight-gradient kernel, not under the gathers
        gpu_timer_begin(TMR_SPMATMUL_BW);
        if (z.Xp && z.tc_transform) GCNK_CHECK(gcnk_dense_transform_bw_tc(z.Xp, z.ld, n_loc, F, z.dxw, W1.grad, H, drop ? z.keep0 : nullptr, scale, z.bw_ws, z.bw_ws_bytes, st));
        else if (z.Xp) GCNK_CHECK(gcnk_dense_transform_bw_ld(z.Xp, z.ld, n_loc, F, z.dxw, W1.grad, H, drop ? z.keep0 : nullptr, scale, z.bw_ws, z.bw_ws_bytes, st));
        else GCNK_CHECK(gcnk_spmm_bw(sp, d_feature_value, z.dxw, W1.grad, H, drop ? z.keep0 : nullptr, scale, st));
        gpu_timer_end(TMR_SPMATMUL_BW);
    }
    finish_pass(training, seq, slot);
}

// The host sync of the enqueued pass(es) and their scalars.
std::pair<float, float> GCN::fused_collect(int slot, bool sync) {
    Fused &z = *fz;
    if (sync) {
        if (z.p2p) GCNK_CHECK(gcnk_memcpy_d2h(z.h_err, z.d_err, sizeof(int), z.stream));
        const int *d_async = nullptr;
        GCNK_CHECK(gcnk_async_error_flag(&d_async));
        GCNK_CHECK(gcnk_memcpy_d2h(z.h_async, d_async, sizeof(int), z.stream));
        GCNK_CHECK(gcnk_stream_sync(z.stream));
        if (*z.h_async) { fprintf(stderr, "GCN: a TMA pipeline kernel timed out on an mbarrier (code %d)\n", *z.h_async); exit(EXIT_FAILURE); }
        if (z.p2p && *z.h_err) { fprintf(stderr, "GCN: a peer rank did not reach the exchange within GCN_PEER_TIMEOUT_S\n"); exit(EXIT_FAILURE); }
        gpu_timer_resolve();
        if (z.sumsq_pending) { z.sumsq = *z.h_sumsq; z.sumsq_pending = false; }
    }
    const float *red = z.h_red + 4 * slot;
    last_count = (int)red[1];
    last_wrong = (int)red[2];
    const float mean_loss = (z.seq_used[slot] ? red[3] : red[0]) / (float)last_count;   // count == 0 -> NaN, as the reference
    const float sumsq = z.sumsq_used[slot] >= 0.f ? z.sumsq_used[slot] : z.sumsq;
    const float l2 = params.weight_decay * sumsq / 2;
    return {mean_loss + l2, float(last_count - last_wrong) / last_count};
}

std::pair<float, float> GCN::fused_pass(int current_split, bool training) {
    fused_enqueue(current_split, training, 0);
    return fused_collect(0, true);
}

// train_epoch() followed by eval(split) with ONE host synchronisation: the eval pass is enqueued behind the
// training pass before anything is read back (what GCN::run does every epoch, gcn.cpp:136-138).
void GCN::epoch(int eval_split, float *train_loss, float *train_acc, float *eval_loss, float *eval_acc) {
    if (plan_ != PLAN_FUSED) {
        std::tie(*train_loss, *train_acc) = train_epoch();
        std::tie(*eval_loss, *eval_acc) = eval(eval_split);
        return;
    }
    timer_start(TMR_HOST_ENQUEUE);
    fused_enqueue(1, true, 0);
    fused_enqueue(eval_split, false, 1);
    timer_stop(TMR_HOST_ENQUEUE);
    std::tie(*train_loss, *train_acc) = fused_collect(0, true);
    train_count = last_count; train_wrong = last_wrong;
    std::tie(*eval_loss, *eval_acc) = fused_collect(1, false);
}

// -------------------------------------------------------------------------------- the loop ----
std::pair<float, float> GCN::train_epoch() {
    if (plan_ == PLAN_FUSED) return fused_pass(1, true);
    set_input();
    set_truth(1);
    for (auto m : modules) m->forward(true);
    const float train_loss = loss + get_l2_penalty();
    const float train_acc = get_accuracy();
    for (int i = (int)modules.size() - 1; i >= 0; i--) modules[i]->backward();
    optimizer.step();
    gpu_timer_resolve();
    return {train_loss, train_acc};
}

std::pair<float, float> GCN::eval(int current_split) {
    if (plan_ == PLAN_FUSED) return fused_pass(current_split, false);
    set_input();
    set_truth(current_split);
    for (auto m : modules) m->forward(false);
    const float test_loss = loss + get_l2_penalty();
    const float test_acc = get_accuracy();
    gpu_timer_resolve();
    return {test_loss, test_acc};
}

void GCN::run() {
    // line formats: gcn.cpp:139,147,152,157
    int epoch = 1;
    std::vector<float> loss_history;
    for (; epoch <= params.epochs; epoch++) {
        float train_loss, train_acc, val_loss, val_acc;
        timer_start(TMR_TRAIN);
        this->epoch(2, &train_loss, &train_acc, &val_loss, &val_acc);
        const float dt = timer_stop(TMR_TRAIN);
        epochs_run = epoch;
        if (!quiet_)
            printf("epoch=%d train_loss=%.5f train_acc=%.5f val_loss=%.5f val_acc=%.5f time=%.5f\n", epoch, train_loss, train_acc,
                   val_loss, val_acc, dt);
        loss_history.push_back(val_loss);
        if (params.early_stopping > 0 && epoch >= params.early_stopping) {
            float recent_loss = 0.0;
            for (int i = epoch - params.early_stopping; i < epoch; i++) recent_loss += loss_history[i];
            if (val_loss > recent_loss / params.early_stopping) {
                if (!quiet_) printf("Early stopping...\n");
                break;
            }
        }
    }
    if (!quiet_) printf("total training time=%.5f\n", timer_total(TMR_TRAIN));
    float test_loss, test_acc;
    timer_start(TMR_TEST);
    std::tie(test_loss, test_acc) = eval(3);
    const float dt = timer_stop(TMR_TEST);
    if (!quiet_) printf("test_loss=%.5f test_acc=%.5f time=%.5f\n", test_loss, test_acc, dt);
}

// ------------------------------------------------------------------------------ inspection ----
long GCN::var_size(int idx) const {
    if (dist.world > 1 && idx != 2 && idx != 5) return 0;      // a partitioned run exposes the (replicated) weights only
    if (fz && fz->wide && idx != 2 && idx != 5 && idx != 3 && idx != 6) return 0;   // wide plan: weights, layer-1 output, logits
    const long N = params.num_nodes, F = params.input_dim, H = params.hidden_dim, C = params.output_dim;
    switch (idx) {
    case 0: return (long)data->feature_index.indices.size();
    case 1: case 3: return N * H;
    case 2: return F * H;
    case 4: return plan_ == PLAN_MODULES ? N * C : 0;
    case 5: return H * C;
    case 6: return N * C;
    }
    return 0;
}

void GCN::get_var(int idx, bool grad, float *h_out) {
    const long size = var_size(idx);
    if (size == 0) return;
    auto d2h = [&](const float *d, long count) {
        GCNK_CHECK(gcnk_memcpy_d2h(h_out, d, sizeof(float) * (size_t)count, nullptr));
        GCNK_CHECK(gcnk_stream_sync(nullptr));
    };
    if (plan_ == PLAN_MODULES) {
        const Variable &v = variables[idx];
        if (grad && !v.grad) { memset(h_out, 0, sizeof(float) * (size_t)size); return; }
        d2h(grad ? v.grad : v.data, size);
        return;
    }
    Fused &z = *fz;
    const int N = params.num_nodes, H = params.hidden_dim, C = params.output_dim;
    if (idx == 2 || idx == 5) { d2h(grad ? variables[idx].grad : variables[idx].data, size); return; }
    if (z.wide) {
        // layer-1 output (data: after ReLU/dropout; grad: the masked gradient that reached it) and the logits of the rows the
        // last pass aggregated (the other rows keep whatever an earlier pass left there)
        if (idx == 3) { d2h(grad ? z.dH1 : z.H1, size); return; }
        if (grad) { memset(h_out, 0, sizeof(float) * (size_t)size); return; }
        float *d_tmp = nullptr;
        GCNK_CHECK(gcnk_malloc((void **)&d_tmp, sizeof(float) * (size_t)size));
        GCNK_CHECK(gcnk_unpad_cols(z.logits, d_tmp, N, C, z.Cp, nullptr));
        d2h(d_tmp, size);
        GCNK_CHECK(gcnk_free(d_tmp));
        return;
    }
    if (idx == 0) { if (grad) memset(h_out, 0, sizeof(float) * (size_t)size); else d2h(d_feature_value, size); return; }
    if (idx == 6) {
        // the logits are never stored by the fused plan: recompute them from the last pass's P with the CURRENT W2
        if (grad) { memset(h_out, 0, sizeof(float) * (size_t)size); return; }
        float *d_logits = nullptr;
        GCNK_CHECK(gcnk_malloc((void **)&d_logits, sizeof(float) * (size_t)size));
        GCNK_CHECK(gcnk_gather_plain(graph_handle(), z.h1_s, z.P, H, nullptr));   // the passes aggregate only their split's rows
        GCNK_CHECK(gcnk_layer2_fused(z.P, variables[5].data, d_split, d_label, 0, N, H, C, 0, 0, nullptr, nullptr, nullptr, d_logits,
                                     z.d_result, z.ws, z.ws_bytes, nullptr));
        d2h(d_logits, size);
        GCNK_CHECK(gcnk_free(d_logits));
        return;
    }
    // 1 and 3: stored pre-scaled by d^-1/2; undo it on the host (inspection only)
    const float *src = idx == 1 ? (grad ? z.dxw : z.xw_s) : (grad ? z.Gm : z.h1_s);
    d2h(src, size);
    if (!(idx == 1 && grad)) {
        const float *d_dinv = nullptr;
        GCNK_CHECK(gcnk_graph_dinv(graph_handle(), &d_dinv));
        std::vector<float> dinv((size_t)N);
        GCNK_CHECK(gcnk_memcpy_d2h(dinv.data(), d_dinv, sizeof(float) * (size_t)N, nullptr));
        GCNK_CHECK(gcnk_stream_sync(nullptr));
        for (int i = 0; i < N; i++)
            for (int j = 0; j < H; j++) h_out[(size_t)i * H + j] /= dinv[i];
    }
}
