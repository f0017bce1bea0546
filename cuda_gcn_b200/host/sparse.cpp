#include "sparse.h"

#include <cstdio>

#include "check.h"

void SparseIndex::print() const {
    printf("---sparse index info--\nindptr: ");
    for (int v : indptr) printf("%d ", v);
    printf("\nindices: ");
    for (int v : indices) printf("%d ", v);
    printf("\n");
}

SparseIndex::~SparseIndex() { release_device(); }

void SparseIndex::release_device() {
    if (graph_) gcnk_graph_destroy(graph_);
    if (spmat_) gcnk_spmat_destroy(spmat_);
    if (dev_indptr_) gcnk_free(dev_indptr_);
    if (dev_indices_) gcnk_free(dev_indices_);
    graph_ = nullptr; spmat_ = nullptr; dev_indptr_ = nullptr; dev_indices_ = nullptr;
}

void SparseIndex::upload() {
    if (dev_indptr_) return;
    static const int zero = 0;
    const int *ip = indptr.empty() ? &zero : indptr.data();
    const size_t n_ip = indptr.empty() ? 1 : indptr.size();
    GCNK_CHECK(gcnk_malloc((void **)&dev_indptr_, sizeof(int) * n_ip));
    GCNK_CHECK(gcnk_malloc((void **)&dev_indices_, sizeof(int) * (indices.size() + 4)));   // + 4: the gather's int4 index reads round up
    GCNK_CHECK(gcnk_memcpy_h2d(dev_indptr_, ip, sizeof(int) * n_ip, nullptr));
    GCNK_CHECK(gcnk_memcpy_h2d(dev_indices_, indices.data(), sizeof(int) * indices.size(), nullptr));
    GCNK_CHECK(gcnk_stream_sync(nullptr));
}

const int *SparseIndex::d_indptr() { upload(); return dev_indptr_; }
const int *SparseIndex::d_indices() { upload(); return dev_indices_; }

gcnk_graph *SparseIndex::graph() {
    if (!graph_) {
        upload();
        GCNK_CHECK(gcnk_graph_create(&graph_, dev_indptr_, dev_indices_, rows(), nnz(), rows(), nullptr, nullptr));
    }
    return graph_;
}

gcnk_graph *SparseIndex::graph_slice(int n_cols, const float *d_dinv_global) {
    if (!graph_) {
        upload();
        GCNK_CHECK(gcnk_graph_create(&graph_, dev_indptr_, dev_indices_, rows(), nnz(), n_cols, d_dinv_global, nullptr));
    }
    return graph_;
}

gcnk_spmat *SparseIndex::spmat(int m, int n) {
    if (!spmat_) {
        upload();
        GCNK_CHECK(gcnk_spmat_create(&spmat_, dev_indptr_, dev_indices_, m, n, nnz(), nullptr));
    }
    return spmat_;
}
