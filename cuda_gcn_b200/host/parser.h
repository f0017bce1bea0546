// parser.h — the data parser with the reference's public shape (src/common/parser.h:14-28):
// Parser(GCNParams*, GCNData*, graph_name) + bool parse(), reading data/<name>.{graph,split,svmlight}.
// The accepted-input behaviour of src/common/parser.cpp:20-119 is reproduced bit-exactly (implicit
// self loop first in every row, neighbours in file order with duplicates kept, a last line without
// '\n' dropped, 0-based feature keys, input_dim = max key + 1, output_dim = max label + 1); the
// istringstream-per-line tokenisers are replaced by one pass over the file bytes.  Malformed
// feature tokens, which the reference turns into uninitialised garbage, are an
// error here: parse() prints what is wrong and returns false.
#pragma once
#include <cstdint>
#include <string>

#include "gcn.h"

// Binary dataset cache: the parsed GCNData + the three parser-derived GCNParams fields as one file of
// raw little-endian arrays, so a Reddit-size dataset loads at file-system speed instead of re-tokenising ~2 GB of text.
// Parser::parse() uses <root>/<name>.gcnbin when its header records exactly the current size and mtime (ns) of the three
// text files, and writes it after a successful text parse unless $GCN_NO_CACHE is set or set_cache_write(false) (only
// one rank of a multi-GPU launch writes it; every writer uses a temporary name of its own and renames).  The cached
// arrays are exactly the parser's output.  `stamp` = the six numbers of source_stamp(); NULL: do not record / check.
bool source_stamp(const std::string &graph, const std::string &split, const std::string &svmlight, int64_t stamp[6]);
bool save_dataset_cache(const std::string &path, const GCNParams &params, const GCNData &data, const int64_t *stamp = nullptr);
bool load_dataset_cache(const std::string &path, GCNParams *params, GCNData *data, const int64_t *stamp = nullptr);

class Parser {
public:
    // root defaults to the reference's hard-coded "data/" (parser.cpp:12); $GCN_DATA_DIR overrides it
    Parser(GCNParams *gcnParams, GCNData *gcnData, std::string graph_name, std::string root = "");
    bool parse();
private:
    std::string graph_path, split_path, svmlight_path, cache_path;
    GCNParams *gcnParams;
    GCNData *gcnData;
    bool quiet = false, write_cache = true;
    bool parseGraph(const std::string &bytes);
    bool parseNode(const std::string &bytes);
    bool parseSplit(const std::string &bytes);
public:
    void set_quiet(bool q) { quiet = q; }
    void set_cache_write(bool w) { write_cache = w; }
};
