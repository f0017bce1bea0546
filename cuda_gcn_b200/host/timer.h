// timer.h — the reference's timer API (src/common/timer.h:5-26): 13 named accumulators, of which only
// TMR_TRAIN / TMR_TEST feed the CLI output (gcn.cpp:140,152,157).  TMR_TRAIN/TMR_TEST are host
// wall-clock as in the reference.  The per-op slots are fed by CUDA-event pairs recorded around each
// kernel group when GPU timing is enabled (the reference's per-op numbers are launch latency only:
// every cudaDeviceSynchronize in cuda_module.cu is commented out).
#pragma once

typedef enum {
    TMR_TRAIN = 0, TMR_TEST, TMR_MATMUL_FW, TMR_MATMUL_BW, TMR_SPMATMUL_FW, TMR_SPMATMUL_BW,
    TMR_GRAPHSUM_FW, TMR_GRAPHSUM_BW, TMR_LOSS_FW, TMR_RELU_FW, TMR_RELU_BW, TMR_DROPOUT_FW, TMR_DROPOUT_BW,
    // additions of this engine (not in the reference enum)
    TMR_ADAM, TMR_COMM,
    TMR_GATHER_FULL, TMR_GATHER_PART,   // GraphSum gather launches over the whole graph / over a row- or column-subset view
    TMR_HOST_ENQUEUE,                   // host wall-clock spent enqueueing a fused epoch (launch-bound when it nears the step time)
    __NUM_TMR
} timer_instance;

void timer_start(timer_instance t);
float timer_stop(timer_instance t);      // seconds since the matching start; also accumulated
float timer_total(timer_instance t);
int timer_calls(timer_instance t);
void timer_add(timer_instance t, float seconds);
void timer_reset_all();
const char *timer_name(timer_instance t);

// CUDA-event timing of device work on the engine's stream.  begin/end only record events (no sync);
// gpu_timer_resolve() must be called after a stream/device synchronisation and folds the elapsed
// times into the slots.  All three are no-ops while disabled.
void gpu_timer_enable(bool on);
void gpu_timer_enable_mask(unsigned mask);   // bit t set = slot t is timed (an event pair per op is not free: ~6 us)
bool gpu_timer_enabled();
void gpu_timer_set_stream(void *stream);      // the stream begin/end record on (default: the legacy stream); per host thread
void gpu_timer_begin(timer_instance t);
void gpu_timer_end(timer_instance t);
void gpu_timer_resolve();

#define PRINT_TIMER_AVERAGE(T, E) printf(#T " average time: %.3fms\n", timer_total(T) * 1000 / E)
