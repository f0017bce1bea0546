#include "timer.h"

#include <chrono>

namespace {
using clk = std::chrono::steady_clock;
struct Slot { clk::time_point t0; double sum = 0; };
Slot g_slots[__NUM_TMR];
}  // namespace

void timer_start(timer_instance t) { g_slots[t].t0 = clk::now(); }

float timer_stop(timer_instance t) {
    const double dt = std::chrono::duration<double>(clk::now() - g_slots[t].t0).count();
    g_slots[t].sum += dt;
    return (float)dt;
}

float timer_total(timer_instance t) { return (float)g_slots[t].sum; }
void timer_add(timer_instance t, float seconds) { g_slots[t].sum += seconds; }

const char *timer_name(timer_instance t) {
    static const char *names[__NUM_TMR] = {"train", "test", "matmul_fw", "matmul_bw", "spmatmul_fw", "spmatmul_bw",
                                           "graphsum_fw", "graphsum_bw", "loss_fw", "relu_fw", "relu_bw", "dropout_fw", "dropout_bw"};
    return t < __NUM_TMR ? names[t] : "?";
}
