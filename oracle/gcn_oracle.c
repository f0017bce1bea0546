/* oracle/gcn_oracle.c — TEST INFRASTRUCTURE (the parity checker), never shipped, never a fallback.
 *
 * Plain-C restatement of the reference's sequential CPU algorithm (see gcn_oracle.h).  Arithmetic is
 * written so that every intermediate has the same type, rounding and evaluation order as the
 * reference's C++ expressions compiled with `-O3 -std=c++11` for baseline x86-64 (SSE2, no FMA):
 * fp32 products and sums stay fp32, expressions that the reference promotes to double through a
 * `1.0`/`0.5` literal are promoted here as well.  Build with -ffp-contract=off (oracle/Makefile).
 *
 * Pinned: bit-exact against oracle/_ref/libgcnref.so (the unmodified reference) in
 * tests/test_oracle_vs_ref.py, and against tests/golden/ produced by that reference build.
 */
#include "gcn_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ RNG (src/seq/rand.cpp) ---- */

static uint64_t g_state[2];
#define GCNO_RAND_MAX 0x7fffffff /* rand.h:6 — an int, so float/int divisions convert it to 2147483648.0f */

/* rand.cpp:6-15.  The reference seeds glibc's rand() from time(NULL); the seed is a parameter here. */
void gcno_init_rand_state(long seed) {
    srand((unsigned)seed);
    int x = 0, y = 0;
    while (x == 0 || y == 0) {
        x = rand();
        y = rand();
    }
    g_state[0] = (uint64_t)x;
    g_state[1] = (uint64_t)y;
}

void gcno_set_rand_state(uint64_t s0, uint64_t s1) { g_state[0] = s0; g_state[1] = s1; }
void gcno_get_rand_state(uint64_t *out2) { out2[0] = g_state[0]; out2[1] = g_state[1]; }

/* rand.cpp:17-28: xorshift128+ with shifts 23/17/26, output masked to 31 bits */
uint32_t gcno_rand(void) {
    uint64_t t = g_state[0];
    const uint64_t s = g_state[1];
    g_state[0] = s;
    t ^= t << 23;
    t ^= t >> 17;
    t ^= s ^ (s >> 26);
    g_state[1] = t;
    return (uint32_t)((t + s) & 0x7fffffffu);
}

/* ------------------------------------------------------- Variable::glorot (variable.cpp:11-18) ---- */

void gcno_glorot(float *w, int in_size, int out_size) {
    const float range = sqrtf(6.0f / (float)(in_size + out_size));
    const long n = (long)in_size * out_size;
    for (long i = 0; i < n; i++) {
        /* float(RAND()) / MY_RAND_MAX - 0.5 : fp32 quotient, then a double subtraction, narrowed to fp32 */
        const float q = (float)gcno_rand() / (float)GCNO_RAND_MAX;
        const float r = (float)((double)q - 0.5);
        w[i] = r * range * 2;
    }
}

/* ----------------------------------------------------------------- Matmul (module.cpp:11-42) ---- */

void gcno_matmul_fw(const float *a, const float *b, float *c, int m, int n, int p) {
    memset(c, 0, sizeof(float) * (size_t)m * p);                       /* c->zero() :14 */
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            const float aij = a[(size_t)i * n + j];
            for (int k = 0; k < p; k++) c[(size_t)i * p + k] += aij * b[(size_t)j * p + k];
        }
}

void gcno_matmul_bw(const float *a, const float *b, const float *c_grad, float *a_grad, float *b_grad,
                    int m, int n, int p) {
    memset(b_grad, 0, sizeof(float) * (size_t)n * p);                  /* b->zero_grad() :28 */
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            float tmp = 0;
            const float aij = a[(size_t)i * n + j];
            for (int k = 0; k < p; k++) {
                const float g = c_grad[(size_t)i * p + k];
                tmp += g * b[(size_t)j * p + k];
                b_grad[(size_t)j * p + k] += g * aij;
            }
            a_grad[(size_t)i * n + j] = tmp;                             /* assigned, not accumulated :37 */
        }
}

/* ----------------------------------------------------------- SparseMatmul (module.cpp:47-77) ---- */

void gcno_spmm_fw(const int *indptr, const int *indices, const float *values, const float *b, float *c,
                  int m, int n, int p) {
    (void)n;
    memset(c, 0, sizeof(float) * (size_t)m * p);
    for (int i = 0; i < m; i++)
        for (int jj = indptr[i]; jj < indptr[i + 1]; jj++) {
            const int j = indices[jj];
            const float x = values[jj];
            for (int k = 0; k < p; k++) c[(size_t)i * p + k] += x * b[(size_t)j * p + k];
        }
}

void gcno_spmm_bw(const int *indptr, const int *indices, const float *values, const float *c_grad,
                  float *b_grad, int m, int n, int p) {
    memset(b_grad, 0, sizeof(float) * (size_t)n * p);
    for (int i = 0; i < m; i++)
        for (int jj = indptr[i]; jj < indptr[i + 1]; jj++) {
            const int j = indices[jj];
            const float x = values[jj];
            for (int k = 0; k < p; k++) b_grad[(size_t)j * p + k] += c_grad[(size_t)i * p + k] * x;
        }
}

/* -------------------------------------------------------------- GraphSum (module.cpp:83-119) ---- */

/* Forward and backward are the same loop over (data) resp. (grad) buffers: out[src] += coef*in[dst],
 * never the transpose (module.cpp:95 comment).  coef: int degree product -> float -> sqrtf ->
 * double reciprocal -> float (module.cpp:91-93). */
void gcno_graphsum(const int *indptr, const int *indices, int n, int dim, const float *in, float *out) {
    memset(out, 0, sizeof(float) * (size_t)n * dim);
    for (int src = 0; src < n; src++) {
        const int deg_src = indptr[src + 1] - indptr[src];
        float *o = out + (size_t)src * dim;
        for (int i = indptr[src]; i < indptr[src + 1]; i++) {
            const int dst = indices[i];
            const int prod = deg_src * (indptr[dst + 1] - indptr[dst]);   /* 32-bit int product, as the reference */
            const float coef = (float)(1.0 / (double)sqrtf((float)prod));
            const float *r = in + (size_t)dst * dim;
            for (int j = 0; j < dim; j++) o[j] += coef * r[j];
        }
    }
}

/* ----------------------------------------------------- CrossEntropyLoss (module.cpp:124-161) ---- */

float gcno_cross_entropy(float *logits, const int *truth, float *grad, int n, int num_classes, int training) {
    float total_loss = 0;
    int count = 0;
    const size_t size = (size_t)n * num_classes;
    if (training) memset(grad, 0, sizeof(float) * size);               /* zero_grad :129 */
    for (int i = 0; i < n; i++) {
        if (truth[i] < 0) continue;
        count++;
        float *logit = logits + (size_t)i * num_classes;
        float max_logit = -1e30f, sum_exp = 0;                           /* -1e30 narrowed to float :135 */
        for (int j = 0; j < num_classes; j++) max_logit = fmaxf(max_logit, logit[j]);
        for (int j = 0; j < num_classes; j++) {
            logit[j] -= max_logit;                                       /* in place :140 */
            sum_exp += expf(logit[j]);
        }
        total_loss += logf(sum_exp) - logit[truth[i]];
        if (training) {
            float *g = grad + (size_t)i * num_classes;
            for (int j = 0; j < num_classes; j++) g[j] = expf(logit[j]) / sum_exp;
            g[truth[i]] = (float)((double)g[truth[i]] - 1.0);
        }
    }
    if (training)
        for (size_t i = 0; i < size; i++) grad[i] /= (float)count;       /* whole array :156-158 */
    return total_loss / (float)count;                                    /* count==0 -> NaN, as the reference */
}

/* --------------------------------------------------------------------- ReLU (module.cpp:175-194) ---- */

void gcno_relu_fw(float *x, unsigned char *mask, int n, int training) {
    for (int i = 0; i < n; i++) {
        const int keep = x[i] > 0;
        if (training) mask[i] = (unsigned char)keep;
        if (!keep) x[i] = 0;
    }
}

void gcno_relu_bw(float *grad, const unsigned char *mask, int n) {
    for (int i = 0; i < n; i++)
        if (!mask[i]) grad[i] = 0;
}

/* ------------------------------------------------------------------ Dropout (module.cpp:207-233) ---- */

void gcno_dropout_fw(float *x, int *mask, int n, float p, int training) {
    if (!training) return;                                               /* no RNG draws in eval :208 */
    const int threshold = (int)(p * (float)GCNO_RAND_MAX);
    const float scale = 1 / (1 - p);
    for (int i = 0; i < n; i++) {
        const int keep = (int)gcno_rand() >= threshold;
        x[i] *= keep ? scale : 0;
        if (mask) mask[i] = keep;
    }
}

void gcno_dropout_bw(float *grad, const int *mask, int n, float p) {
    if (!mask) return;
    const float scale = 1 / (1 - p);
    for (int i = 0; i < n; i++) grad[i] *= mask[i] ? scale : 0;
}

/* ------------------------------------------------------------------------- Adam (optim.cpp:6-37) ---- */

struct gcno_adam {
    int nvars, step_count;
    int *sizes, *decay;
    float **m, **v;
    float lr, beta1, beta2, eps, weight_decay;
};

gcno_adam *gcno_adam_create(int nvars, const int *sizes, const int *decay, float lr, float beta1, float beta2,
                            float eps, float weight_decay) {
    gcno_adam *o = (gcno_adam *)calloc(1, sizeof *o);
    o->nvars = nvars;
    o->sizes = (int *)malloc(sizeof(int) * nvars);
    o->decay = (int *)malloc(sizeof(int) * nvars);
    o->m = (float **)malloc(sizeof(float *) * nvars);
    o->v = (float **)malloc(sizeof(float *) * nvars);
    for (int i = 0; i < nvars; i++) {
        o->sizes[i] = sizes[i];
        o->decay[i] = decay[i];
        o->m[i] = (float *)calloc((size_t)sizes[i], sizeof(float));     /* m, v start at 0 :11 */
        o->v[i] = (float *)calloc((size_t)sizes[i], sizeof(float));
    }
    o->lr = lr; o->beta1 = beta1; o->beta2 = beta2; o->eps = eps; o->weight_decay = weight_decay;
    return o;
}

void gcno_adam_step(gcno_adam *o, float **data, float **grad) {
    o->step_count++;
    const float sc = (float)o->step_count;
    const float step_size = o->lr * sqrtf(1 - powf(o->beta2, sc)) / (1 - powf(o->beta1, sc));
    for (int k = 0; k < o->nvars; k++) {
        float *m = o->m[k], *v = o->v[k], *w = data[k];
        const float *gr = grad[k];
        for (int i = 0; i < o->sizes[k]; i++) {
            float g = gr[i];
            if (o->decay[k]) g += o->weight_decay * w[i];
            /* beta*m is an fp32 product; (1.0 - beta)*g is double; the sum is double, stored as fp32 */
            m[i] = (float)((double)(o->beta1 * m[i]) + (1.0 - (double)o->beta1) * (double)g);
            v[i] = (float)((double)(o->beta2 * v[i]) + (1.0 - (double)o->beta2) * (double)g * (double)g);
            w[i] -= step_size * m[i] / (sqrtf(v[i]) + o->eps);
        }
    }
}

void gcno_adam_destroy(gcno_adam *o) {
    if (!o) return;
    for (int i = 0; i < o->nvars; i++) { free(o->m[i]); free(o->v[i]); }
    free(o->m); free(o->v); free(o->sizes); free(o->decay); free(o);
}

/* ----------------------------------------------------------------- GCN helpers (gcn.cpp:78-105) ---- */

void gcno_set_truth(int *truth, const int *split, const int *label, int n, int current_split) {
    for (int i = 0; i < n; i++) truth[i] = split[i] == current_split ? label[i] : -1;
}

/* wrong iff some logit is strictly greater than the truth logit (ties count as correct) */
float gcno_accuracy(const float *logits, const int *truth, int n, int num_classes, int *wrong_out, int *total_out) {
    int wrong = 0, total = 0;
    for (int i = 0; i < n; i++) {
        if (truth[i] < 0) continue;
        total++;
        const float *row = logits + (size_t)i * num_classes;
        const float t = row[truth[i]];
        for (int j = 0; j < num_classes; j++)
            if (row[j] > t) { wrong++; break; }
    }
    if (wrong_out) *wrong_out = wrong;
    if (total_out) *total_out = total;
    return (float)(total - wrong) / (float)total;
}

float gcno_l2_penalty(const float *w, int size, float weight_decay) {
    float l2 = 0;
    for (int i = 0; i < size; i++) l2 += w[i] * w[i];
    return weight_decay * l2 / 2;
}

/* ------------------------------------------------------------------ Parser (parser.cpp:20-119) ---- */

typedef struct { int *p; long n, cap; } ivec;
typedef struct { float *p; long n, cap; } fvec;
static void ipush(ivec *v, int x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = (int *)realloc(v->p, sizeof(int) * v->cap); }
    v->p[v->n++] = x;
}
static void fpush(fvec *v, float x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = (float *)realloc(v->p, sizeof(float) * v->cap); }
    v->p[v->n++] = x;
}

static char *slurp(const char *dir, const char *name, const char *ext, long *len) {
    char path[4096];
    snprintf(path, sizeof path, "%s/data/%s.%s", dir, name, ext);      /* root = "data/" :12 */
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)n + 1);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return NULL; }
    fclose(f);
    buf[n] = 0;
    *len = n;
    return buf;
}

/* getline(); if (eof) break;  => only '\n'-terminated lines are seen (:27-28,:62-63,:99-100).
 * Returns the next terminated line (NUL-terminated in place) or NULL. */
static char *next_line(char **cur, char *end) {
    char *s = *cur;
    if (s >= end) return NULL;
    char *nl = (char *)memchr(s, '\n', (size_t)(end - s));
    if (!nl) return NULL;                                                /* unterminated last line is dropped */
    *nl = 0;
    *cur = nl + 1;
    return s;
}

static int is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }

/* `ss >> int`: skip whitespace, optional sign, digits; fails (returns 0) if no digits follow. */
static int read_int(char **s, int *out) {
    char *p = *s;
    while (is_ws(*p)) p++;
    char *e;
    long v = strtol(p, &e, 10);
    if (e == p) return 0;
    *out = (int)v;
    *s = e;
    return 1;
}

int gcno_parse(gcno_data *d, const char *dir, const char *name) {
    memset(d, 0, sizeof *d);
    long glen, slen, vlen;
    char *gbuf = slurp(dir, name, "graph", &glen);
    char *sbuf = slurp(dir, name, "split", &slen);
    char *vbuf = slurp(dir, name, "svmlight", &vlen);
    if (!gbuf || !sbuf || !vbuf) { free(gbuf); free(sbuf); free(vbuf); return 0; }   /* isValidInput :48-50 */

    /* parseGraph :20-46 — implicit self loop first, then the neighbours in file order, no dedup */
    ivec gptr = {0}, gidx = {0};
    ipush(&gptr, 0);
    int node = 0;
    char *cur = gbuf, *line;
    while ((line = next_line(&cur, gbuf + glen))) {
        ipush(&gidx, node);
        ipush(&gptr, gptr.p[gptr.n - 1] + 1);
        node++;
        int nb;
        while (read_int(&line, &nb)) { ipush(&gidx, nb); gptr.p[gptr.n - 1] += 1; }
    }
    d->num_nodes = node;

    /* parseNode :52-92 */
    ivec fptr = {0}, fidx = {0}, lab = {0};
    fvec fval = {0};
    ipush(&fptr, 0);
    int max_idx = 0, max_label = 0;
    cur = vbuf;
    while ((line = next_line(&cur, vbuf + vlen))) {
        ipush(&fptr, fptr.p[fptr.n - 1]);
        int label = -1;
        char *p = line;
        while (is_ws(*p)) p++;
        if (*p == 0) { ipush(&lab, -1); continue; }                      /* blank line: label stays -1 :67-70 */
        if (!read_int(&p, &label)) { ipush(&lab, 0); continue; }         /* C++11 num_get stores 0 on failure */
        ipush(&lab, label);
        if (label > max_label) max_label = label;
        for (;;) {
            while (is_ws(*p)) p++;
            if (*p == 0) break;
            char *tok = p;
            while (*p && !is_ws(*p)) p++;
            char saved = *p;
            *p = 0;
            /* kv_ss >> k >> col >> v :82 */
            int k = 0;
            float v = 0;
            char *q = tok;
            if (read_int(&q, &k)) {
                while (is_ws(*q)) q++;
                if (*q) { q++; char *e; v = strtof(q, &e); }
            }
            *p = saved;
            fpush(&fval, v);
            ipush(&fidx, k);
            fptr.p[fptr.n - 1] += 1;
            if (k > max_idx) max_idx = k;
        }
    }
    d->input_dim = max_idx + 1;
    d->output_dim = max_label + 1;

    /* parseSplit :94-103 — std::stoi per line */
    ivec spl = {0};
    cur = sbuf;
    while ((line = next_line(&cur, sbuf + slen))) ipush(&spl, (int)strtol(line, NULL, 10));

    d->graph_indptr = gptr.p; d->graph_indices = gidx.p; d->graph_nnz = gidx.n;
    d->feature_indptr = fptr.p; d->feature_indices = fidx.p; d->feature_value = fval.p; d->feature_nnz = fidx.n;
    d->label = lab.p; d->n_label = lab.n;
    d->split = spl.p; d->n_split = spl.n;
    free(gbuf); free(sbuf); free(vbuf);
    return 1;
}

void gcno_data_free(gcno_data *d) {
    free(d->graph_indptr); free(d->graph_indices); free(d->feature_indptr); free(d->feature_indices);
    free(d->feature_value); free(d->label); free(d->split);
    memset(d, 0, sizeof *d);
}

/* ------------------------------------------------------------ the model and loop (gcn.cpp) ---- */

struct gcno_gcn {
    const gcno_data *d;
    gcno_hparams hp;
    int N, F, H, C;
    /* variables in construction order gcn.cpp:21-53 */
    float *input;                 /* V0 [nnzX], no grad */
    float *xw, *xw_g;             /* V1 [N*H] */
    float *w1, *w1_g;             /* V2 [F*H] */
    float *h1, *h1_g;             /* V3 [N*H] */
    float *hw, *hw_g;             /* V4 [N*C] */
    float *w2, *w2_g;             /* V5 [H*C] */
    float *out, *out_g;           /* V6 [N*C] */
    unsigned char *relu_mask;
    int *drop_mask;               /* layer-1 dropout only; the input dropout keeps none (module.cpp:199-200) */
    int *truth;
    gcno_adam *opt;
    float loss;
};

gcno_hparams gcno_default_hparams(void) {
    gcno_hparams hp = {16, 0.5f, 0.01f, 5e-4f, 100, 0};                 /* gcn.cpp:10 */
    return hp;
}

static float *falloc(size_t n) { return (float *)calloc(n ? n : 1, sizeof(float)); }

gcno_gcn *gcno_gcn_create(const gcno_data *d, gcno_hparams hp, long seed) {
    gcno_gcn *g = (gcno_gcn *)calloc(1, sizeof *g);
    g->d = d; g->hp = hp;
    g->N = d->num_nodes; g->F = d->input_dim; g->H = hp.hidden_dim; g->C = d->output_dim;
    const size_t N = g->N, F = g->F, H = g->H, C = g->C;
    gcno_init_rand_state(seed);                                          /* :14 */
    g->input = falloc((size_t)d->feature_nnz);
    g->xw = falloc(N * H); g->xw_g = falloc(N * H);
    g->w1 = falloc(F * H); g->w1_g = falloc(F * H);
    gcno_glorot(g->w1, g->F, g->H);                                      /* :30 — W1 draws first */
    g->h1 = falloc(N * H); g->h1_g = falloc(N * H);
    g->hw = falloc(N * C); g->hw_g = falloc(N * C);
    g->w2 = falloc(H * C); g->w2_g = falloc(H * C);
    gcno_glorot(g->w2, g->H, g->C);                                      /* :49 — then W2 */
    g->out = falloc(N * C); g->out_g = falloc(N * C);
    g->relu_mask = (unsigned char *)calloc(N * H + 1, 1);
    g->drop_mask = (int *)calloc(N * H + 1, sizeof(int));
    g->truth = (int *)calloc(N ? N : 1, sizeof(int));
    int sizes[2] = {(int)(F * H), (int)(H * C)}, decay[2] = {1, 0};      /* :65 */
    g->opt = gcno_adam_create(2, sizes, decay, hp.learning_rate, 0.9f, 0.999f, 1e-8f, hp.weight_decay);
    return g;
}

void gcno_gcn_destroy(gcno_gcn *g) {
    if (!g) return;
    free(g->input); free(g->xw); free(g->xw_g); free(g->w1); free(g->w1_g); free(g->h1); free(g->h1_g);
    free(g->hw); free(g->hw_g); free(g->w2); free(g->w2_g); free(g->out); free(g->out_g);
    free(g->relu_mask); free(g->drop_mask); free(g->truth);
    gcno_adam_destroy(g->opt);
    free(g);
}

static void forward(gcno_gcn *g, int training) {
    const gcno_data *d = g->d;
    const int NH = g->N * g->H;
    gcno_dropout_fw(g->input, NULL, (int)d->feature_nnz, g->hp.dropout, training);                    /* M0 */
    gcno_spmm_fw(d->feature_indptr, d->feature_indices, g->input, g->w1, g->xw, g->N, g->F, g->H);  /* M1 */
    gcno_graphsum(d->graph_indptr, d->graph_indices, g->N, g->H, g->xw, g->h1);                     /* M2 */
    gcno_relu_fw(g->h1, g->relu_mask, NH, training);                                                 /* M3 */
    gcno_dropout_fw(g->h1, g->drop_mask, NH, g->hp.dropout, training);                               /* M4 */
    gcno_matmul_fw(g->h1, g->w2, g->hw, g->N, g->H, g->C);                                           /* M5 */
    gcno_graphsum(d->graph_indptr, d->graph_indices, g->N, g->C, g->hw, g->out);                    /* M6 */
    g->loss = gcno_cross_entropy(g->out, g->truth, g->out_g, g->N, g->C, training);                  /* M7 */
}

static void backward(gcno_gcn *g) {
    const gcno_data *d = g->d;
    const int NH = g->N * g->H;
    gcno_graphsum(d->graph_indptr, d->graph_indices, g->N, g->C, g->out_g, g->hw_g);                /* M6 bw */
    gcno_matmul_bw(g->h1, g->w2, g->hw_g, g->h1_g, g->w2_g, g->N, g->H, g->C);                       /* M5 bw */
    gcno_dropout_bw(g->h1_g, g->drop_mask, NH, g->hp.dropout);                                       /* M4 bw */
    gcno_relu_bw(g->h1_g, g->relu_mask, NH);                                                         /* M3 bw */
    gcno_graphsum(d->graph_indptr, d->graph_indices, g->N, g->H, g->h1_g, g->xw_g);                 /* M2 bw */
    gcno_spmm_bw(d->feature_indptr, d->feature_indices, g->input, g->xw_g, g->w1_g, g->N, g->F, g->H); /* M1 bw */
}

void gcno_gcn_train_epoch(gcno_gcn *g, float *loss, float *acc) {
    memcpy(g->input, g->d->feature_value, sizeof(float) * (size_t)g->d->feature_nnz);               /* set_input :73-76 */
    gcno_set_truth(g->truth, g->d->split, g->d->label, g->N, 1);
    forward(g, 1);
    *loss = g->loss + gcno_l2_penalty(g->w1, g->F * g->H, g->hp.weight_decay);
    *acc = gcno_accuracy(g->out, g->truth, g->N, g->C, NULL, NULL);
    backward(g);
    float *data[2] = {g->w1, g->w2}, *grad[2] = {g->w1_g, g->w2_g};
    gcno_adam_step(g->opt, data, grad);
}

void gcno_gcn_eval(gcno_gcn *g, int split, float *loss, float *acc) {
    memcpy(g->input, g->d->feature_value, sizeof(float) * (size_t)g->d->feature_nnz);
    gcno_set_truth(g->truth, g->d->split, g->d->label, g->N, split);
    forward(g, 0);
    *loss = g->loss + gcno_l2_penalty(g->w1, g->F * g->H, g->hp.weight_decay);
    *acc = gcno_accuracy(g->out, g->truth, g->N, g->C, NULL, NULL);
}

static float *var_ptr(gcno_gcn *g, int idx, int grad, long *size) {
    const long N = g->N, F = g->F, H = g->H, C = g->C;
    switch (idx) {
    case 0: *size = g->d->feature_nnz; return grad ? NULL : g->input;
    case 1: *size = N * H; return grad ? g->xw_g : g->xw;
    case 2: *size = F * H; return grad ? g->w1_g : g->w1;
    case 3: *size = N * H; return grad ? g->h1_g : g->h1;
    case 4: *size = N * C; return grad ? g->hw_g : g->hw;
    case 5: *size = H * C; return grad ? g->w2_g : g->w2;
    case 6: *size = N * C; return grad ? g->out_g : g->out;
    }
    *size = 0;
    return NULL;
}

long gcno_gcn_var_size(gcno_gcn *g, int idx) { long s; var_ptr(g, idx, 0, &s); return s; }
void gcno_gcn_get_var(gcno_gcn *g, int idx, int grad, float *out) {
    long s;
    float *p = var_ptr(g, idx, grad, &s);
    if (p) memcpy(out, p, sizeof(float) * (size_t)s);
}

#include <time.h>
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* run(): gcn.cpp:130-158 */
int gcno_gcn_run(gcno_gcn *g) {
    int epoch = 1, done = 0;
    float *hist = (float *)malloc(sizeof(float) * (size_t)(g->hp.epochs > 0 ? g->hp.epochs : 1));
    double total = 0;
    for (; epoch <= g->hp.epochs; epoch++) {
        float tl, ta, vl, va;
        double t0 = now_s();
        gcno_gcn_train_epoch(g, &tl, &ta);
        gcno_gcn_eval(g, 2, &vl, &va);
        double dt = now_s() - t0;
        total += dt;
        printf("epoch=%d train_loss=%.5f train_acc=%.5f val_loss=%.5f val_acc=%.5f time=%.5f\n", epoch, tl, ta, vl, va, dt);
        hist[epoch - 1] = vl;
        done = epoch;
        if (g->hp.early_stopping > 0 && epoch >= g->hp.early_stopping) {
            float recent = 0.0f;
            for (int i = epoch - g->hp.early_stopping; i < epoch; i++) recent += hist[i];
            if (vl > recent / (float)g->hp.early_stopping) { printf("Early stopping...\n"); break; }
        }
    }
    printf("total training time=%.5f\n", total);
    float tl, ta;
    double t0 = now_s();
    gcno_gcn_eval(g, 3, &tl, &ta);
    printf("test_loss=%.5f test_acc=%.5f time=%.5f\n", tl, ta, now_s() - t0);
    fflush(stdout);
    free(hist);
    return done;
}
