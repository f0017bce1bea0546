// check.h — the reference's error convention on top of the C ABI's return codes: print and exit
// (CUDA_CHECK in src/cuda/cuda_kernel.cuh:11-18).  There is no fallback path: any failure of the CUDA
// library is fatal.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "gcnk.h"

#define GCNK_CHECK(call)                                                                          \
    do {                                                                                          \
        const int _rc = (call);                                                                   \
        if (_rc != GCNK_OK) {                                                                     \
            fprintf(stderr, "CUDA_ASSERT: %s (%d) %s %d\n", gcnk_last_error(), _rc, __FILE__, __LINE__); \
            exit(_rc > 0 ? _rc : 1);                                                              \
        }                                                                                         \
    } while (0)
