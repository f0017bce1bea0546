#include "sparse.h"

#include <cstdio>

void SparseIndex::print() const {
    printf("---sparse index info--\nindptr: ");
    for (int v : indptr) printf("%d ", v);
    printf("\nindices: ");
    for (int v : indices) printf("%d ", v);
    printf("\n");
}
