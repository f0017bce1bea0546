// runtime.cu — device/memory/stream/event entry points of the C ABI (include/gcnk.h, first two groups).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gcnk {

std::atomic<int64_t> g_launches{0};
static thread_local char t_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof t_error, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return (int)e;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// One int per device that the mbarrier/TMA/tcgen05 kernels raise when a wait times out (a protocol error must not
// hang the GPU, but it must not pass silently either): callers poll it with gcnk_async_error.
int *async_err_flag() {
    static int *d_err[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!d_err[dev]) {
        if (cudaMalloc(&d_err[dev], sizeof(int)) != cudaSuccess) return nullptr;
        cudaMemset(d_err[dev], 0, sizeof(int));
    }
    return d_err[dev];
}

}  // namespace gcnk

using namespace gcnk;

extern "C" {

int gcnk_version(void) { return 100; }
const char *gcnk_last_error(void) { return t_error; }

int gcnk_async_error_flag(const int **d_flag) {
    GCNK_REQUIRE(d_flag, "null");
    *d_flag = async_err_flag();
    if (!*d_flag) return cuda_fail(cudaGetLastError(), "async error flag", __FILE__, __LINE__);
    return GCNK_OK;
}

int gcnk_async_error(gcnk_stream_t stream) {
    int *d = async_err_flag(), h = 0;
    if (!d) return cuda_fail(cudaGetLastError(), "async error flag", __FILE__, __LINE__);
    GCNK_CUDA(cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    GCNK_CUDA(cudaStreamSynchronize(S(stream)));
    if (h) {
        GCNK_CUDA(cudaMemsetAsync(d, 0, sizeof(int), S(stream)));
        set_error("a TMA / tcgen05 pipeline kernel timed out on an mbarrier (code %d): its output is invalid", h);
        return GCNK_EASYNC;
    }
    return GCNK_OK;
}

int gcnk_device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        if (count) *count = 0;
        cudaGetLastError();
        set_error("no CUDA device (%s)", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
        return GCNK_ENODEVICE;
    }
    if (count) *count = n;
    return GCNK_OK;
}

int gcnk_set_device(int device) { GCNK_CUDA(cudaSetDevice(device)); return GCNK_OK; }

int gcnk_device_info(int device, int *sms, int *cc_major, int *cc_minor, size_t *total_mem, int *l2_bytes) {
    cudaDeviceProp p;
    GCNK_CUDA(cudaGetDeviceProperties(&p, device));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    return GCNK_OK;
}

int64_t gcnk_launch_count(void) { return g_launches.load(); }

int gcnk_malloc(void **ptr, size_t bytes) {
    GCNK_REQUIRE(ptr, "null out pointer");
    *ptr = nullptr;
    if (bytes == 0) bytes = 16;
    GCNK_CUDA(cudaMalloc(ptr, bytes));
    return GCNK_OK;
}
int gcnk_free(void *ptr) { if (ptr) GCNK_CUDA(cudaFree(ptr)); return GCNK_OK; }
int gcnk_malloc_host(void **ptr, size_t bytes) {
    GCNK_REQUIRE(ptr, "null out pointer");
    GCNK_CUDA(cudaMallocHost(ptr, bytes ? bytes : 16));
    return GCNK_OK;
}
int gcnk_free_host(void *ptr) { if (ptr) GCNK_CUDA(cudaFreeHost(ptr)); return GCNK_OK; }

int gcnk_memcpy_h2d(void *dst, const void *src, size_t bytes, gcnk_stream_t s) {
    if (bytes) GCNK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, S(s)));
    return GCNK_OK;
}
int gcnk_memcpy_d2h(void *dst, const void *src, size_t bytes, gcnk_stream_t s) {
    if (bytes) GCNK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S(s)));
    return GCNK_OK;
}
int gcnk_memcpy_d2d(void *dst, const void *src, size_t bytes, gcnk_stream_t s) {
    if (bytes) GCNK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, S(s)));
    return GCNK_OK;
}
int gcnk_memset(void *dst, int value, size_t bytes, gcnk_stream_t s) {
    if (bytes) GCNK_CUDA(cudaMemsetAsync(dst, value, bytes, S(s)));
    return GCNK_OK;
}

int gcnk_stream_create(gcnk_stream_t *stream) {
    cudaStream_t s;
    GCNK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return GCNK_OK;
}
int gcnk_stream_create_low_priority(gcnk_stream_t *stream) {
    int least = 0, greatest = 0;
    GCNK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    cudaStream_t s;
    GCNK_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, least));
    *stream = s;
    return GCNK_OK;
}
int gcnk_stream_destroy(gcnk_stream_t s) { if (s) GCNK_CUDA(cudaStreamDestroy(S(s))); return GCNK_OK; }
int gcnk_stream_sync(gcnk_stream_t s) { GCNK_CUDA(cudaStreamSynchronize(S(s))); return GCNK_OK; }
int gcnk_device_sync(void) { GCNK_CUDA(cudaDeviceSynchronize()); return GCNK_OK; }

int gcnk_event_create(void **ev) {
    cudaEvent_t e;
    GCNK_CUDA(cudaEventCreate(&e));
    *ev = e;
    return GCNK_OK;
}
int gcnk_event_destroy(void *ev) { if (ev) GCNK_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return GCNK_OK; }
int gcnk_event_record(void *ev, gcnk_stream_t s) { GCNK_CUDA(cudaEventRecord((cudaEvent_t)ev, S(s))); return GCNK_OK; }
int gcnk_event_sync(void *ev) { GCNK_CUDA(cudaEventSynchronize((cudaEvent_t)ev)); return GCNK_OK; }
int gcnk_stream_wait_event(gcnk_stream_t s, void *ev) { GCNK_CUDA(cudaStreamWaitEvent(S(s), (cudaEvent_t)ev, 0)); return GCNK_OK; }
int gcnk_event_elapsed_ms(void *a, void *b, float *ms) {
    GCNK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
    return GCNK_OK;
}

// 256 MiB > the 126 MB L2: one streaming write evicts whatever the previous iteration left resident
static __global__ void flush_kernel(float4 *p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
int gcnk_flush_l2(gcnk_stream_t s) {
    static void *buf[64] = {nullptr};
    const size_t bytes = 256u << 20;
    int dev = 0;
    GCNK_CUDA(cudaGetDevice(&dev));
    if (!buf[dev]) GCNK_CUDA(cudaMalloc(&buf[dev], bytes));
    flush_kernel<<<sm_count() * 8, 256, 0, S(s)>>>((float4 *)buf[dev], bytes / sizeof(float4));
    GCNK_LAUNCHED();
    return GCNK_OK;
}

}  // extern "C"
