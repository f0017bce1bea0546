#include "rand.h"

#include <cstdlib>
#include <ctime>

#include "check.h"

// one stream per host thread: a process drives one GPU per thread at most (rand.cpp:5 keeps one global state)
static thread_local gcnk_rng *g_rng = nullptr;

gcnk_rng *global_rng() {
    if (!g_rng) GCNK_CHECK(gcnk_rng_create(&g_rng, 1, 2));
    return g_rng;
}

void init_rand_state(long seed) { GCNK_CHECK(gcnk_rng_seed(global_rng(), seed)); }

void init_rand_state() {
    const char *s = getenv("GCN_SEED");
    init_rand_state(s && *s ? atol(s) : (long)time(NULL));
}

uint32_t gcn_rand() {
    uint32_t v;
    GCNK_CHECK(gcnk_rng_next_host(global_rng(), &v, 1));
    return v;
}
