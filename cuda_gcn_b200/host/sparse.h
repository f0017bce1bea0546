// sparse.h — host CSR index, the same public shape as the reference's SparseIndex
// (src/seq/sparse.h:12-17): used for both the graph and the feature matrix.
#pragma once
#include <vector>

class SparseIndex {
public:
    std::vector<int> indices;   // column ids, len nnz
    std::vector<int> indptr;    // row offsets, len nrow + 1
    void print() const;
};
