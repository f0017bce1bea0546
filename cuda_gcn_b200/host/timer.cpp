#include "timer.h"

#include <chrono>
#include <vector>

#include "check.h"

namespace {
using clk = std::chrono::steady_clock;
struct Slot { clk::time_point t0; double sum = 0; int calls = 0; };
thread_local Slot g_slots[__NUM_TMR];

struct Pending { timer_instance t; void *start, *stop; };
thread_local bool g_gpu_on = false;
thread_local unsigned g_gpu_mask = 0xffffffffu;
thread_local std::vector<Pending> g_pending;
thread_local std::vector<void *> g_free_events;
thread_local void *g_open[__NUM_TMR] = {nullptr};
thread_local void *g_stream = nullptr;        // the stream the event pairs are recorded on

void *take_event() {
    if (!g_free_events.empty()) { void *e = g_free_events.back(); g_free_events.pop_back(); return e; }
    void *e = nullptr;
    GCNK_CHECK(gcnk_event_create(&e));
    return e;
}
}  // namespace

void timer_start(timer_instance t) { g_slots[t].t0 = clk::now(); }

float timer_stop(timer_instance t) {
    const double dt = std::chrono::duration<double>(clk::now() - g_slots[t].t0).count();
    g_slots[t].sum += dt;
    g_slots[t].calls++;
    return (float)dt;
}

float timer_total(timer_instance t) { return (float)g_slots[t].sum; }
int timer_calls(timer_instance t) { return g_slots[t].calls; }
void timer_add(timer_instance t, float seconds) { g_slots[t].sum += seconds; g_slots[t].calls++; }
void timer_reset_all() { for (auto &s : g_slots) { s.sum = 0; s.calls = 0; } }

const char *timer_name(timer_instance t) {
    static const char *names[__NUM_TMR] = {"train", "test", "matmul_fw", "matmul_bw", "spmatmul_fw", "spmatmul_bw",
                                           "graphsum_fw", "graphsum_bw", "loss_fw", "relu_fw", "relu_bw", "dropout_fw",
                                           "dropout_bw", "adam", "comm", "gather_full", "gather_part", "host_enqueue"};
    return t < __NUM_TMR ? names[t] : "?";
}

void gpu_timer_enable(bool on) { g_gpu_on = on; g_gpu_mask = 0xffffffffu; }
void gpu_timer_enable_mask(unsigned mask) { g_gpu_on = mask != 0; g_gpu_mask = mask; }
bool gpu_timer_enabled() { return g_gpu_on; }
void gpu_timer_set_stream(void *stream) { g_stream = stream; }

void gpu_timer_begin(timer_instance t) {
    if (!g_gpu_on || !((g_gpu_mask >> t) & 1u)) return;
    void *e = take_event();
    GCNK_CHECK(gcnk_event_record(e, g_stream));
    g_open[t] = e;
}

void gpu_timer_end(timer_instance t) {
    if (!g_gpu_on || !g_open[t]) return;
    void *e = take_event();
    GCNK_CHECK(gcnk_event_record(e, g_stream));
    g_pending.push_back({t, g_open[t], e});
    g_open[t] = nullptr;
}

void gpu_timer_resolve() {
    for (const Pending &p : g_pending) {
        float ms = 0;
        GCNK_CHECK(gcnk_event_sync(p.stop));
        GCNK_CHECK(gcnk_event_elapsed_ms(p.start, p.stop, &ms));
        timer_add(p.t, ms * 1e-3f);
        g_free_events.push_back(p.start);
        g_free_events.push_back(p.stop);
    }
    g_pending.clear();
}
