// rand.h — the process-wide xorshift128+ stream (reference src/seq/rand.{h,cpp}).  The state lives in
// a gcnk_rng so that host draws (Glorot init) and the device dropout masks consume ONE stream in the
// reference's order: W1 draws, W2 draws, then per training pass nnz(X) + N*H draws.
#pragma once
#include <cstdint>

#include "gcnk.h"

#define MY_RAND_MAX 0x7fffffff

// Seeds like init_rand_state() (rand.cpp:6-15): srand(seed); rand(); rand().  The seed is $GCN_SEED when
// set, else time(NULL) as in the reference.
void init_rand_state();
void init_rand_state(long seed);
gcnk_rng *global_rng();
uint32_t gcn_rand();            // one host draw, RAND() in the reference
#define RAND() gcn_rand()
