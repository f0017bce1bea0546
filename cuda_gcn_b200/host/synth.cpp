#include "synth.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// counter-based uniform generator: value i of stream s does not depend on how work is split over threads
inline uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
inline uint64_t draw(uint64_t seed, uint64_t stream, uint64_t i) { return mix64(mix64(seed * 0x100000001b3ull + stream) ^ (i * 0xd6e8feb86659fd93ull)); }
inline double unit(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)

template <typename Fn>
void parallel_for(int64_t n, Fn fn) {
    unsigned t = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < 1 << 16) t = 1;
    std::vector<std::thread> th;
    const int64_t chunk = (n + t - 1) / t;
    for (unsigned k = 0; k < t; k++) {
        const int64_t lo = k * chunk, hi = std::min<int64_t>(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=] { fn(lo, hi); });
    }
    for (auto &x : th) x.join();
}

}  // namespace

bool synth_preset(const char *name, double scale, SynthSpec *out) {
    SynthSpec s{};
    s.alpha = 1.0; s.isolated_nodes = 0;
    if (!strcmp(name, "cora"))          s = {2708, 5278, 1.2, 1433, 18, 7, 140.0 / 2708, 500.0 / 2708, 1000.0 / 2708, 1, 0};
    else if (!strcmp(name, "citeseer")) s = {3327, 4552, 1.2, 3703, 32, 6, 120.0 / 3327, 500.0 / 3327, 1000.0 / 3327, 2, 48};
    else if (!strcmp(name, "pubmed"))   s = {19717, 44324, 1.3, 500, 50, 3, 60.0 / 19717, 500.0 / 19717, 1000.0 / 19717, 3, 0};
    else if (!strcmp(name, "reddit"))   s = {232965, 58450000, 1.55, 602, 0, 41, 0.66, 0.10, 0.24, 4, 0};
    else if (!strcmp(name, "products")) s = {2449029, 61859140, 1.45, 100, 0, 47, 0.08, 0.02, 0.90, 5, 0};
    else return false;
    if (scale > 0 && scale != 1.0) {
        // fewer nodes, same mean degree and the same feature/class widths
        s.num_nodes = std::max(64, (int)(s.num_nodes * scale));
        s.undirected_edges = std::max<int64_t>(64, (int64_t)(s.undirected_edges * scale));
        s.isolated_nodes = (int)(s.isolated_nodes * scale);
    }
    *out = s;
    return true;
}

bool synth_generate(const SynthSpec &spec, GCNParams *params, GCNData *data) {
    const int N = spec.num_nodes, F = spec.input_dim, C = spec.output_dim;
    const int64_t E = spec.undirected_edges;
    const int live = std::max(1, N - spec.isolated_nodes);

    // ---- node permutation: hubs (small a) must not be contiguous in id space
    std::vector<int> perm((size_t)N);
    for (int i = 0; i < N; i++) perm[i] = i;
    for (int i = N - 1; i > 0; i--) std::swap(perm[i], perm[(size_t)(draw(spec.seed, 1, (uint64_t)i) % (uint64_t)(i + 1))]);

    // ---- endpoints
    std::vector<int> ea((size_t)E), eb((size_t)E);
    parallel_for(E, [&](int64_t lo, int64_t hi) {
        for (int64_t e = lo; e < hi; e++) {
            const double u = unit(draw(spec.seed, 2, (uint64_t)e)), w = unit(draw(spec.seed, 3, (uint64_t)e));
            int a = (int)(live * std::pow(u, spec.alpha)), b = (int)(live * w);
            if (a >= live) a = live - 1;
            if (b >= live) b = live - 1;
            // homophily: 60% of the edges join two nodes of the same class (class of raw id r is r % C)
            if (unit(draw(spec.seed, 10, (uint64_t)e)) < 0.6) {
                b = b - b % C + a % C;
                if (b >= live) b -= C;
                if (b < 0) b = a;
            }
            ea[e] = perm[a]; eb[e] = perm[b];
        }
    });

    // ---- bucket both directions by source row, sort + unique every row
    std::vector<int64_t> start((size_t)N + 1, 0);
    for (int64_t e = 0; e < E; e++)
        if (ea[e] != eb[e]) { start[ea[e] + 1]++; start[eb[e] + 1]++; }
    for (int i = 0; i < N; i++) start[i + 1] += start[i];
    std::vector<int> nb((size_t)start[N]);
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t e = 0; e < E; e++)
            if (ea[e] != eb[e]) { nb[cur[ea[e]]++] = eb[e]; nb[cur[eb[e]]++] = ea[e]; }
    }
    std::vector<int>().swap(ea);
    std::vector<int>().swap(eb);
    std::vector<int> uniq((size_t)N);
    parallel_for(N, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            int *b = nb.data() + start[i], *e = nb.data() + start[i + 1];
            std::sort(b, e);
            uniq[i] = (int)(std::unique(b, e) - b);
        }
    });
    std::vector<int> &indptr = data->graph.indptr, &indices = data->graph.indices;
    indptr.assign((size_t)N + 1, 0);
    int64_t total = 0;
    int max_deg = 0;
    for (int i = 0; i < N; i++) {
        total += 1 + uniq[i];
        max_deg = std::max(max_deg, 1 + uniq[i]);
        if (total > INT32_MAX) { fprintf(stderr, "synth: graph nnz exceeds int32\n"); return false; }
        indptr[i + 1] = (int)total;
    }
    if (max_deg > 46340) {
        fprintf(stderr, "synth: max degree %d > 46340 would overflow the reference's int32 degree product\n", max_deg);
        return false;
    }
    indices.resize((size_t)total);
    parallel_for(N, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) {
            int *o = indices.data() + indptr[i];
            *o++ = (int)i;                                                 // the parser's implicit self loop comes first
            std::copy(nb.data() + start[i], nb.data() + start[i] + uniq[i], o);
        }
    });
    std::vector<int>().swap(nb);

    // ---- labels (class of raw id r is r % C; ids are permuted) and split
    data->label.resize((size_t)N);
    data->split.resize((size_t)N);
    for (int r = 0; r < N; r++) data->label[perm[r]] = r % C;
    for (int i = 0; i < N; i++) {
        const double u = unit(draw(spec.seed, 9, (uint64_t)i));
        data->split[i] = u < spec.train_frac ? 1 : u < spec.train_frac + spec.val_frac ? 2
                         : u < spec.train_frac + spec.val_frac + spec.test_frac ? 3 : 0;
    }

    // ---- features
    std::vector<int> &fptr = data->feature_index.indptr, &fidx = data->feature_index.indices;
    std::vector<float> &fval = data->feature_value;
    fptr.assign((size_t)N + 1, 0);
    if (spec.feature_nnz_per_row <= 0) {
        if ((int64_t)N * F > INT32_MAX) { fprintf(stderr, "synth: dense feature nnz exceeds int32\n"); return false; }
        fidx.resize((size_t)N * F);
        fval.resize((size_t)N * F);
        for (int i = 0; i <= N; i++) fptr[i] = i * F;
        parallel_for(N, [&](int64_t lo, int64_t hi) {
            for (int64_t i = lo; i < hi; i++)
                for (int f = 0; f < F; f += 2) {
                    // Box-Muller: N(0,1), like StandardScaler output (reddit_preprocess.py:71-77)
                    const uint64_t id = (uint64_t)i * F + f;
                    const double u1 = 1.0 - unit(draw(spec.seed, 4, id)), u2 = unit(draw(spec.seed, 5, id));
                    const double r = std::sqrt(-2.0 * std::log(u1)), th = 6.283185307179586 * u2;
                    // columns congruent to the node's class carry a +0.5 mean shift (a learnable signal)
                    const int cls = data->label[i];
                    fidx[(size_t)i * F + f] = f;
                    fval[(size_t)i * F + f] = (float)(r * std::cos(th) + (f % C == cls ? 0.5 : 0.0));
                    if (f + 1 < F) {
                        fidx[(size_t)i * F + f + 1] = f + 1;
                        fval[(size_t)i * F + f + 1] = (float)(r * std::sin(th) + ((f + 1) % C == cls ? 0.5 : 0.0));
                    }
                }
        });
    } else {
        std::vector<std::vector<int>> rows((size_t)N);
        parallel_for(N, [&](int64_t lo, int64_t hi) {
            for (int64_t i = lo; i < hi; i++) {
                const int want = std::max(1, std::min(F, spec.feature_nnz_per_row / 2 + (int)(draw(spec.seed, 6, (uint64_t)i) % (uint64_t)(spec.feature_nnz_per_row + 1))));
                std::vector<int> &r = rows[i];
                // half of the words come from the band of the vocabulary that belongs to the node's class
                const int cls = data->label[i], band = std::max(1, F / C);
                for (int k = 0; k < want; k++) {
                    const uint64_t rr = draw(spec.seed, 7, (uint64_t)i * 4096 + k);
                    r.push_back((rr >> 40) & 1 ? std::min(F - 1, cls * band + (int)(rr % (uint64_t)band)) : (int)(rr % (uint64_t)F));
                }
                std::sort(r.begin(), r.end());
                r.erase(std::unique(r.begin(), r.end()), r.end());
            }
        });
        rows[N - 1].push_back(F - 1);                                      // max key + 1 == F (parser.cpp:90)
        std::sort(rows[N - 1].begin(), rows[N - 1].end());
        rows[N - 1].erase(std::unique(rows[N - 1].begin(), rows[N - 1].end()), rows[N - 1].end());
        for (int i = 0; i < N; i++) fptr[i + 1] = fptr[i] + (int)rows[i].size();
        fidx.resize((size_t)fptr[N]);
        fval.resize((size_t)fptr[N]);
        for (int i = 0; i < N; i++) {
            const float v = 1.0f / (float)rows[i].size();                  // row-normalised bag of words
            for (size_t k = 0; k < rows[i].size(); k++) { fidx[fptr[i] + k] = rows[i][k]; fval[fptr[i] + k] = v; }
        }
    }

    params->num_nodes = N;
    params->input_dim = F;
    params->output_dim = C;
    return true;
}
