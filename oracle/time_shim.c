/* oracle/time_shim.c — TEST INFRASTRUCTURE.  LD_PRELOAD this to pin time(NULL), the reference's only
 * seed source (rand.cpp:7), to $GCN_SEED so two gcn-seq runs are reproducible.  Unset => real time. */
#define _GNU_SOURCE
#include <time.h>
#include <stdlib.h>

time_t time(time_t *t) {
    const char *s = getenv("GCN_SEED");
    time_t v;
    if (s && *s) v = (time_t)atol(s);
    else { struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts); v = ts.tv_sec; }
    if (t) *t = v;
    return v;
}
