// timer.h — wall-clock accumulators behind the reference's timer API (src/common/timer.h:5-26).
// Only TMR_TRAIN / TMR_TEST feed the CLI output (gcn.cpp:140,152,157); the per-op slots are kept for
// API compatibility and are fed by the engine's CUDA-event timers when profiling is on.
#pragma once

typedef enum {
    TMR_TRAIN = 0, TMR_TEST, TMR_MATMUL_FW, TMR_MATMUL_BW, TMR_SPMATMUL_FW, TMR_SPMATMUL_BW,
    TMR_GRAPHSUM_FW, TMR_GRAPHSUM_BW, TMR_LOSS_FW, TMR_RELU_FW, TMR_RELU_BW, TMR_DROPOUT_FW, TMR_DROPOUT_BW,
    __NUM_TMR
} timer_instance;

void timer_start(timer_instance t);
float timer_stop(timer_instance t);      // seconds since the matching start; also accumulated
float timer_total(timer_instance t);
void timer_add(timer_instance t, float seconds);
const char *timer_name(timer_instance t);

#define PRINT_TIMER_AVERAGE(T, E) printf(#T " average time: %.3fms\n", timer_total(T) * 1000 / E)
