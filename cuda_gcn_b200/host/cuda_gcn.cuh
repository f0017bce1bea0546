// cuda_gcn.cuh — name shim for code written against the reference's GPU engine: its main.cpp includes
// "cuda_gcn.cuh" and runs `CUDAGCN cuda_gcn(params, &data); cuda_gcn.run();` when __NVCC__ is defined
// (src/main.cpp:9-11,40-42; class declared in src/cuda/cuda_gcn.cuh:30-33).  Here the GPU engine IS `GCN`
// (gcn.h), so the UNMODIFIED reference main.cpp builds against this directory with plain g++:
//   g++ -std=c++17 -D__NVCC__ -I<repo>/include -I<repo>/cuda_gcn_b200/host <reference>/src/main.cpp \
//       -L<repo>/cuda_gcn_b200 -lgcnhost -lgcnk -o gcn-cuda
// (oracle/Makefile builds exactly that as oracle/_ref/gcn-cuda-refmain; tests/test_integration.py runs it.)
#pragma once
#include "gcn.h"

using CUDAGCN = GCN;
