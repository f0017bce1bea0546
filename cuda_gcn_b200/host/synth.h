// synth.h — seeded synthetic datasets of the BASELINE shapes, generated in memory as a
// GCNData: the real cora/citeseer/pubmed/Reddit/ogbn-products files are not available offline, and
// the reference's text parser needs minutes at Reddit size.  The same arrays feed this engine and the
// CPU checker, so parity does not depend on the generator — only the workload shape does.
#pragma once
#include <cstdint>

#include "gcn.h"

struct SynthSpec {
    int num_nodes;            // N
    int64_t undirected_edges; // pairs drawn before symmetrisation / de-duplication
    double alpha;             // power-law skew: endpoint a = floor(N * u^alpha), b = floor(N * u'); 1.0 = uniform
    int input_dim;            // F
    int feature_nnz_per_row;  // 0 => dense rows (all F columns stored, N(0,1) values, Reddit-like)
    int output_dim;           // C
    double train_frac, val_frac, test_frac;
    uint64_t seed;
    int isolated_nodes;       // nodes that keep only the self loop (citeseer has some)
};

// Fills data (graph CSR with the self loop first then sorted unique neighbours, feature CSR, labels,
// split) and params->{num_nodes,input_dim,output_dim}.  Returns false (with a message on stderr) if
// the maximum degree exceeds 46,340 — the reference's int32 deg*deg product would overflow
// (module.cpp:92).
bool synth_generate(const SynthSpec &spec, GCNParams *params, GCNData *data);

// named presets: cora, citeseer, pubmed, reddit, products, optionally scaled down
bool synth_preset(const char *name, double scale, SynthSpec *out);
