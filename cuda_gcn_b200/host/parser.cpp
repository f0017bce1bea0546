#include "parser.h"

#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <iostream>

#include <sys/stat.h>

namespace {

bool slurp(const std::string &path, std::string &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(len > 0 ? (size_t)len : 0);
    const size_t got = len > 0 ? fread(&out[0], 1, (size_t)len, f) : 0;
    fclose(f);
    return got == out.size();
}

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }

// `stream >> int`: skip whitespace, optional sign, decimal digits; fails (consuming nothing useful)
// when no digit follows or the value does not fit an int.  [p, end) never crosses a newline.
inline bool scan_int(const char *&p, const char *end, int &value) {
    while (p < end && is_space(*p)) p++;
    const char *q = p;
    bool neg = false;
    if (q < end && (*q == '+' || *q == '-')) { neg = *q == '-'; q++; }
    if (q >= end || *q < '0' || *q > '9') return false;
    long long v = 0;
    bool overflow = false;
    while (q < end && *q >= '0' && *q <= '9') {
        v = v * 10 + (*q - '0');
        if (v > (long long)INT_MAX + 1) overflow = true, v = (long long)INT_MAX + 1;
        q++;
    }
    p = q;
    if (neg) v = -v;
    if (overflow || v > INT_MAX || v < INT_MIN) return false;
    value = (int)v;
    return true;
}

// iterate over the '\n'-terminated lines of a buffer; a trailing fragment without '\n' is NOT a line
// (getline hits EOF and the reference breaks before using it, parser.cpp:27-28,62-63,99-100)
struct Lines {
    const char *cur, *end;
    explicit Lines(const std::string &s) : cur(s.data()), end(s.data() + s.size()) {}
    bool next(const char *&lo, const char *&hi) {
        const char *p = cur;
        while (p < end && *p != '\n') p++;
        if (p >= end) return false;
        lo = cur; hi = p; cur = p + 1;
        return true;
    }
};

}  // namespace

Parser::Parser(GCNParams *gcnParams_, GCNData *gcnData_, std::string graph_name, std::string root)
    : gcnParams(gcnParams_), gcnData(gcnData_) {
    if (root.empty()) {
        const char *env = getenv("GCN_DATA_DIR");
        root = env && *env ? env : "data/";
    }
    if (root.back() != '/') root += '/';
    graph_path = root + graph_name + ".graph";
    split_path = root + graph_name + ".split";
    svmlight_path = root + graph_name + ".svmlight";
    cache_path = root + graph_name + ".gcnbin";
}

// ------------------------------------------------------------------------------ binary cache ----
namespace {
const char CACHE_MAGIC[8] = {'G', 'C', 'N', 'B', 'I', 'N', '0', '2'};

template <typename T>
bool write_vec(FILE *f, const std::vector<T> &v) {
    const uint64_t n = v.size();
    return fwrite(&n, sizeof n, 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <typename T>
bool read_vec(FILE *f, std::vector<T> &v) {
    uint64_t n = 0;
    if (fread(&n, sizeof n, 1, f) != 1 || n > (1ull << 33)) return false;
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
}  // namespace

// {size, mtime (ns)} of the three text files: the cache is valid only for exactly these files
bool source_stamp(const std::string &graph, const std::string &split, const std::string &svm, int64_t stamp[6]) {
    const std::string *files[3] = {&graph, &split, &svm};
    for (int i = 0; i < 3; i++) {
        struct stat st;
        if (stat(files[i]->c_str(), &st) != 0) return false;
        stamp[2 * i] = (int64_t)st.st_size;
        stamp[2 * i + 1] = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    }
    return true;
}

bool save_dataset_cache(const std::string &path, const GCNParams &params, const GCNData &data, const int64_t *stamp) {
    // a name of its own per writer: concurrent processes (GCN_GPUS=N forks one per GPU) never share a half-written file
    const std::string tmp = path + ".tmp." + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    const int32_t dims[3] = {params.num_nodes, params.input_dim, params.output_dim};
    const int64_t zero[6] = {0, 0, 0, 0, 0, 0};
    bool ok = fwrite(CACHE_MAGIC, 8, 1, f) == 1 && fwrite(dims, sizeof dims, 1, f) == 1 &&
              fwrite(stamp ? stamp : zero, sizeof zero, 1, f) == 1 && write_vec(f, data.graph.indptr) &&
              write_vec(f, data.graph.indices) && write_vec(f, data.feature_index.indptr) && write_vec(f, data.feature_index.indices) &&
              write_vec(f, data.feature_value) && write_vec(f, data.label) && write_vec(f, data.split);
    ok = fclose(f) == 0 && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) { remove(tmp.c_str()); return false; }
    return true;
}

bool load_dataset_cache(const std::string &path, GCNParams *params, GCNData *data, const int64_t *stamp) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[8];
    int32_t dims[3];
    int64_t stored[6];
    GCNData d;
    const bool ok = fread(magic, 8, 1, f) == 1 && !memcmp(magic, CACHE_MAGIC, 8) && fread(dims, sizeof dims, 1, f) == 1 &&
                    fread(stored, sizeof stored, 1, f) == 1 && (!stamp || !memcmp(stored, stamp, sizeof stored)) &&
                    read_vec(f, d.graph.indptr) && read_vec(f, d.graph.indices) && read_vec(f, d.feature_index.indptr) &&
                    read_vec(f, d.feature_index.indices) && read_vec(f, d.feature_value) && read_vec(f, d.label) && read_vec(f, d.split) &&
                    (int)d.graph.indptr.size() == dims[0] + 1 && d.feature_value.size() == d.feature_index.indices.size();
    fclose(f);
    if (!ok) return false;
    data->graph.indptr.swap(d.graph.indptr); data->graph.indices.swap(d.graph.indices);
    data->feature_index.indptr.swap(d.feature_index.indptr); data->feature_index.indices.swap(d.feature_index.indices);
    data->feature_value.swap(d.feature_value); data->label.swap(d.label); data->split.swap(d.split);
    params->num_nodes = dims[0]; params->input_dim = dims[1]; params->output_dim = dims[2];
    return true;
}

bool Parser::parseGraph(const std::string &bytes) {
    std::vector<int> &indptr = gcnData->graph.indptr, &indices = gcnData->graph.indices;
    indptr.clear(); indices.clear();
    indices.reserve(bytes.size() / 4);
    indptr.push_back(0);
    Lines lines(bytes);
    const char *lo, *hi;
    int node = 0;
    while (lines.next(lo, hi)) {
        indices.push_back(node);                       // the implicit self connection comes first
        int nb;
        while (scan_int(lo, hi, nb)) indices.push_back(nb);   // stops at the first non-integer token
        if (indices.size() > (size_t)INT_MAX) { fprintf(stderr, "graph: more than INT_MAX entries\n"); return false; }
        indptr.push_back((int)indices.size());
        node++;
    }
    gcnParams->num_nodes = node;
    return true;
}

bool Parser::parseNode(const std::string &bytes) {
    std::vector<int> &indptr = gcnData->feature_index.indptr, &indices = gcnData->feature_index.indices;
    std::vector<float> &values = gcnData->feature_value;
    std::vector<int> &labels = gcnData->label;
    indptr.clear(); indices.clear(); values.clear(); labels.clear();
    indptr.push_back(0);
    int max_idx = 0, max_label = 0;
    Lines lines(bytes);
    const char *lo, *hi;
    long line_no = 0;
    while (lines.next(lo, hi)) {
        line_no++;
        const char *p = lo;
        while (p < hi && is_space(*p)) p++;
        if (p == hi) {                                  // blank line: label -1, empty feature row
            labels.push_back(-1);
            indptr.push_back((int)indices.size());
            continue;
        }
        int label = 0;
        if (!scan_int(p, hi, label)) {                  // non-numeric label: C++11 extraction stores 0; row skipped
            labels.push_back(0);
            indptr.push_back((int)indices.size());
            continue;
        }
        labels.push_back(label);
        if (label > max_label) max_label = label;
        for (;;) {
            while (p < hi && is_space(*p)) p++;
            if (p == hi) break;
            const char *tok_end = p;
            while (tok_end < hi && !is_space(*tok_end)) tok_end++;
            int k;
            const char *q = p;
            if (!scan_int(q, tok_end, k) || q >= tok_end || *q != ':' || q + 1 >= tok_end) {
                fprintf(stderr, "%s:%ld: malformed feature token '%.*s' (expected key:value)\n", svmlight_path.c_str(),
                        line_no, (int)(tok_end - p), p);
                return false;
            }
            char *after = nullptr;
            const std::string num(q + 1, tok_end);      // NUL-terminated copy for strtof
            const float v = strtof(num.c_str(), &after);
            if (after == num.c_str()) {
                fprintf(stderr, "%s:%ld: malformed feature value in '%.*s'\n", svmlight_path.c_str(), line_no,
                        (int)(tok_end - p), p);
                return false;
            }
            values.push_back(v);
            indices.push_back(k);
            if (k > max_idx) max_idx = k;
            p = tok_end;
        }
        if (indices.size() > (size_t)INT_MAX) { fprintf(stderr, "svmlight: more than INT_MAX entries\n"); return false; }
        indptr.push_back((int)indices.size());
    }
    gcnParams->input_dim = max_idx + 1;
    gcnParams->output_dim = max_label + 1;
    return true;
}

bool Parser::parseSplit(const std::string &bytes) {
    std::vector<int> &split = gcnData->split;
    split.clear();
    Lines lines(bytes);
    const char *lo, *hi;
    long line_no = 0;
    while (lines.next(lo, hi)) {
        line_no++;
        int v;
        if (!scan_int(lo, hi, v)) {                     // std::stoi would throw here (parser.cpp:101)
            fprintf(stderr, "%s:%ld: not an integer\n", split_path.c_str(), line_no);
            return false;
        }
        split.push_back(v);
    }
    return true;
}

bool Parser::parse() {
    const char *nc = getenv("GCN_NO_CACHE");
    const bool use_cache = !(nc && *nc && strcmp(nc, "0"));
    // the cache header records size and mtime (ns) of the three text files it was made from; anything else is stale
    int64_t stamp[6];
    const bool have_text = source_stamp(graph_path, split_path, svmlight_path, stamp);
    if (use_cache && have_text && load_dataset_cache(cache_path, gcnParams, gcnData, stamp)) {
        // the same three lines: callers (and tests) key on them
        if (!quiet) std::cout << "Parse Graph Succeeded." << std::endl << "Parse Node Succeeded." << std::endl << "Parse Split Succeeded." << std::endl;
        return true;
    }
    std::string g, s, v;
    if (!slurp(graph_path, g) || !slurp(split_path, s) || !slurp(svmlight_path, v)) return false;
    if (!parseGraph(g)) return false;
    if (!quiet) std::cout << "Parse Graph Succeeded." << std::endl;
    if (!parseNode(v)) return false;
    if (!quiet) std::cout << "Parse Node Succeeded." << std::endl;
    if (!parseSplit(s)) return false;
    if (!quiet) std::cout << "Parse Split Succeeded." << std::endl;
    if (use_cache && write_cache) save_dataset_cache(cache_path, *gcnParams, *gcnData, stamp);   // best effort: a read-only data/ is not an error
    return true;
}
