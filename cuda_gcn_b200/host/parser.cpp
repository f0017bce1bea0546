#include "parser.h"

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <iostream>

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
    std::string g, s, v;
    if (!slurp(graph_path, g) || !slurp(split_path, s) || !slurp(svmlight_path, v)) return false;
    if (!parseGraph(g)) return false;
    if (!quiet) std::cout << "Parse Graph Succeeded." << std::endl;
    if (!parseNode(v)) return false;
    if (!quiet) std::cout << "Parse Node Succeeded." << std::endl;
    if (!parseSplit(s)) return false;
    if (!quiet) std::cout << "Parse Split Succeeded." << std::endl;
    return true;
}
