// sparse.h — host CSR index with the reference's public shape (src/seq/sparse.h:12-17; used for both
// the graph and the feature matrix), plus a lazily-built device twin (the role of CUDASparseIndex,
// src/cuda/cuda_variable.cuh:22-30): the CSR arrays uploaded once, and the gcnk handles prepared on
// them (gcnk_graph: d^-1/2 + static row schedule; gcnk_spmat: dense detection + CSC view).
#pragma once
#include <cstdint>
#include <vector>

#include "gcnk.h"

class SparseIndex {
public:
    std::vector<int> indices;   // column ids, len nnz
    std::vector<int> indptr;    // row offsets, len nrow + 1
    void print() const;

    SparseIndex() = default;
    ~SparseIndex();
    SparseIndex(const SparseIndex &o) : indices(o.indices), indptr(o.indptr) {}
    SparseIndex &operator=(const SparseIndex &o) { release_device(); indices = o.indices; indptr = o.indptr; return *this; }

    // ---- device twin (not part of the reference API).  Built from the vectors as they are at the
    // first call; call release_device() after editing them.
    int rows() const { return indptr.empty() ? 0 : (int)indptr.size() - 1; }
    int64_t nnz() const { return (int64_t)indices.size(); }
    const int *d_indptr();
    const int *d_indices();
    gcnk_graph *graph();                       // this index as the (square) normalised adjacency
    // this index as a ROW SLICE of an adjacency with n_cols nodes (column ids global); d_dinv_global = d^-1/2 of all nodes
    gcnk_graph *graph_slice(int n_cols, const float *d_dinv_global);
    gcnk_spmat *spmat(int m, int n);           // this index as an m x n sparse feature matrix
    void release_device();

private:
    int *dev_indptr_ = nullptr, *dev_indices_ = nullptr;
    gcnk_graph *graph_ = nullptr;
    gcnk_spmat *spmat_ = nullptr;
    void upload();
};
