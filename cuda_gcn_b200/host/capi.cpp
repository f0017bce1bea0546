// capi.cpp — extern "C" face of the host layer (include/gcn_host.h).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "check.h"
#include "gcn.h"
#include "gcn_host.h"
#include "parser.h"
#include "rand.h"
#include "synth.h"
#include "timer.h"

struct gcnh_data { GCNData d; };
struct gcnh_engine { GCN *g; gcnk_comm *comm = nullptr; float *d_tmp = nullptr; };

static GCNParams to_cpp(const gcnh_params &p) {
    return GCNParams{p.num_nodes, p.input_dim, p.hidden_dim, p.output_dim, p.dropout, p.learning_rate, p.weight_decay, p.epochs, p.early_stopping};
}
static gcnh_params to_c(const GCNParams &p) {
    return gcnh_params{p.num_nodes, p.input_dim, p.hidden_dim, p.output_dim, p.dropout, p.learning_rate, p.weight_decay, p.epochs, p.early_stopping};
}

extern "C" {

gcnh_params gcnh_default_params(void) { return to_c(GCNParams::get_default()); }

gcnh_data *gcnh_data_new(void) { return new gcnh_data; }
void gcnh_data_free(gcnh_data *d) { delete d; }

int gcnh_data_parse(gcnh_data *d, const char *root, const char *name, gcnh_params *params, int quiet) {
    GCNParams p = to_cpp(*params);
    d->d = GCNData();
    Parser parser(&p, &d->d, name, root ? root : "");
    parser.set_quiet(quiet != 0);
    if (!parser.parse()) return 0;
    *params = to_c(p);
    return 1;
}

int gcnh_data_fill(gcnh_data *d, int n, const int *gp, const int *gi, const int *fp, const int *fi, const float *fv,
                   const int *label, const int *split) {
    GCNData &x = d->d;
    x.graph.release_device();
    x.feature_index.release_device();
    x.graph.indptr.assign(gp, gp + n + 1);
    x.graph.indices.assign(gi, gi + gp[n]);
    x.feature_index.indptr.assign(fp, fp + n + 1);
    x.feature_index.indices.assign(fi, fi + fp[n]);
    x.feature_value.assign(fv, fv + fp[n]);
    x.label.assign(label, label + n);
    x.split.assign(split, split + n);
    return 1;
}

int gcnh_data_synth(gcnh_data *d, const char *preset, double scale, uint64_t seed, gcnh_params *params) {
    SynthSpec spec;
    if (!synth_preset(preset, scale, &spec)) { fprintf(stderr, "unknown synthetic preset '%s'\n", preset); return 0; }
    if (seed) spec.seed = seed;
    GCNParams p = to_cpp(*params);
    d->d.graph.release_device();
    d->d.feature_index.release_device();
    if (!synth_generate(spec, &p, &d->d)) return 0;
    *params = to_c(p);
    return 1;
}

gcnh_data *gcnh_data_slice(const gcnh_data *d, int rank, int world, int *row_begin, int *row_end) {
    const int n = d->d.graph.rows();
    std::vector<int> cuts((size_t)world + 1);
    if (gcnk_partition_rows(d->d.graph.indptr.data(), n, world, cuts.data()) != GCNK_OK || rank < 0 || rank >= world) return nullptr;
    gcnh_data *out = new gcnh_data;
    slice_rows(d->d, cuts[rank], cuts[rank + 1] - cuts[rank], out->d);
    if (row_begin) *row_begin = cuts[rank];
    if (row_end) *row_end = cuts[rank + 1];
    return out;
}

void gcnh_data_sizes(const gcnh_data *d, int64_t *s) {
    const GCNData &x = d->d;
    s[0] = x.graph.rows();
    s[1] = x.graph.nnz();
    s[2] = x.feature_index.nnz();
    s[3] = (int64_t)x.label.size();
    s[4] = (int64_t)x.split.size();
    int md = 0;
    for (int i = 0; i < x.graph.rows(); i++) md = std::max(md, x.graph.indptr[i + 1] - x.graph.indptr[i]);
    s[5] = md;
    s[6] = x.feature_index.rows();
}

const int *gcnh_data_graph_indptr(const gcnh_data *d) { return d->d.graph.indptr.data(); }
const int *gcnh_data_graph_indices(const gcnh_data *d) { return d->d.graph.indices.data(); }
const int *gcnh_data_feature_indptr(const gcnh_data *d) { return d->d.feature_index.indptr.data(); }
const int *gcnh_data_feature_indices(const gcnh_data *d) { return d->d.feature_index.indices.data(); }
const float *gcnh_data_feature_value(const gcnh_data *d) { return d->d.feature_value.data(); }
const int *gcnh_data_label(const gcnh_data *d) { return d->d.label.data(); }
const int *gcnh_data_split(const gcnh_data *d) { return d->d.split.data(); }

gcnh_engine *gcnh_engine_create(const gcnh_params *params, gcnh_data *data, long seed, int plan, int device) {
    GCNK_CHECK(gcnk_set_device(device));
    if (seed >= 0) {
        // GCN's constructor seeds from $GCN_SEED or time(NULL) (rand.cpp:6-15); pin it for this construction
        char buf[32];
        snprintf(buf, sizeof buf, "%ld", seed);
        setenv("GCN_SEED", buf, 1);
    }
    gcnh_engine *e = new gcnh_engine;
    e->g = new GCN(to_cpp(*params), &data->d, (GCNPlan)plan, true);
    return e;
}

int gcnh_comm_unique_id(void *id) { return gcnk_comm_unique_id(id) == GCNK_OK; }

gcnh_engine *gcnh_engine_create_dist(const gcnh_params *params, gcnh_data *data, long seed, int device, int rank, int world,
                                     const void *id128) {
    GCNK_CHECK(gcnk_set_device(device));
    if (seed >= 0) {
        char buf[32];
        snprintf(buf, sizeof buf, "%ld", seed);
        setenv("GCN_SEED", buf, 1);
    }
    gcnh_engine *e = new gcnh_engine;
    GCNDist dist;
    dist.rank = rank; dist.world = world;
    if (world > 1) {
        GCNK_CHECK(gcnk_comm_create(&e->comm, id128, rank, world, device));
        dist.comm = e->comm;
    }
    e->g = new GCN(to_cpp(*params), &data->d, PLAN_FUSED, true, dist);
    return e;
}

void gcnh_engine_allreduce_host(gcnh_engine *e, float *h, int count, int op_max) {
    if (!e->comm || count <= 0) return;
    if (!e->d_tmp) GCNK_CHECK(gcnk_malloc((void **)&e->d_tmp, sizeof(float) * 64));
    if (count > 64) count = 64;
    GCNK_CHECK(gcnk_memcpy_h2d(e->d_tmp, h, sizeof(float) * count, nullptr));
    float *bufs[1] = {e->d_tmp};
    const size_t counts[1] = {(size_t)count};
    GCNK_CHECK(gcnk_comm_allreduce(e->comm, bufs, counts, 1, op_max, nullptr));
    GCNK_CHECK(gcnk_memcpy_d2h(h, e->d_tmp, sizeof(float) * count, nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
}

void gcnh_engine_destroy(gcnh_engine *e) {
    if (!e) return;
    delete e->g;
    if (e->d_tmp) gcnk_free(e->d_tmp);
    if (e->comm) gcnk_comm_destroy(e->comm);
    delete e;
}
int gcnh_engine_plan(const gcnh_engine *e) { return (int)e->g->plan(); }

void gcnh_engine_train_epoch(gcnh_engine *e, float *loss, float *acc) {
    auto r = e->g->train_epoch();
    if (loss) *loss = r.first;
    if (acc) *acc = r.second;
}

void gcnh_engine_eval(gcnh_engine *e, int split, float *loss, float *acc) {
    auto r = e->g->eval(split);
    if (loss) *loss = r.first;
    if (acc) *acc = r.second;
}

void gcnh_engine_epoch(gcnh_engine *e, int split, float *tl, float *ta, float *el, float *ea) {
    float a, b, c, d;
    e->g->epoch(split, &a, &b, &c, &d);
    if (tl) *tl = a;
    if (ta) *ta = b;
    if (el) *el = c;
    if (ea) *ea = d;
}

void gcnh_engine_last_counts(const gcnh_engine *e, int *count, int *wrong) {
    if (count) *count = e->g->last_count;
    if (wrong) *wrong = e->g->last_wrong;
}

int gcnh_engine_run(gcnh_engine *e, int quiet) {
    GCN *g = e->g;
    (void)quiet;
    g->run();
    return g->epochs_run;
}

void gcnh_engine_set_input_host(gcnh_engine *e, const float *h) { e->g->set_input_from_host(h); }
void gcnh_engine_epoch_prefetch(gcnh_engine *e, int split, const float *h_next, float *tl, float *ta, float *el, float *ea) {
    float a, b, c, d;
    e->g->epoch_prefetch(split, h_next, &a, &b, &c, &d);
    if (tl) *tl = a;
    if (ta) *ta = b;
    if (el) *el = c;
    if (ea) *ea = d;
}
int64_t gcnh_engine_var_size(const gcnh_engine *e, int idx) { return e->g->var_size(idx); }
void gcnh_engine_get_var(gcnh_engine *e, int idx, int grad, float *out) { e->g->get_var(idx, grad != 0, out); }

void gcnh_timer_enable_gpu(int on) { gpu_timer_enable(on != 0); }
void gcnh_timer_enable_mask(unsigned mask) { gpu_timer_enable_mask(mask); }
int gcnh_timer_slot(const char *name) {
    for (int t = 0; t < __NUM_TMR; t++)
        if (!strcmp(timer_name((timer_instance)t), name)) return t;
    return -1;
}
void gcnh_timer_reset(void) { timer_reset_all(); }
float gcnh_timer_total(int slot) { return slot >= 0 && slot < __NUM_TMR ? timer_total((timer_instance)slot) : 0.f; }
int gcnh_timer_calls(int slot) { return slot >= 0 && slot < __NUM_TMR ? timer_calls((timer_instance)slot) : 0; }
const char *gcnh_timer_name(int slot) { return timer_name((timer_instance)slot); }
int gcnh_timer_count(void) { return __NUM_TMR; }

float *gcnh_alloc_pinned(int64_t n) {
    void *p = nullptr;
    GCNK_CHECK(gcnk_malloc_host(&p, sizeof(float) * (size_t)n));
    return (float *)p;
}
void gcnh_free_pinned(float *p) { gcnk_free_host(p); }

}  // extern "C"
