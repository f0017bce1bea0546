// comm.cu — the exchange steps of the row-partitioned engine over NCCL (NVLink 5 / NVSwitch).
// The reference has no multi-GPU path at all; these entry points are what the partitioned
// plan adds: an all-gather of the GraphSum input with per-rank row counts (the partition is
// nnz-balanced, not row-balanced), and sum / max all-reduces for weight gradients and scalars.
//
// The all-gather is IN PLACE: every rank's producing kernel writes its rows straight into its slice of
// the [N x dim] buffer, and one grouped set of ncclBroadcast calls (one root per rank) fills the rest,
// so there is no staging copy.  The rendezvous (sharing the 128-byte ncclUniqueId) is the caller's job.
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

using namespace gcnk;

struct gcnk_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

#define GCNK_NCCL(expr)                                                                       \
    do {                                                                                      \
        ncclResult_t _r = (expr);                                                             \
        if (_r != ncclSuccess) {                                                              \
            gcnk::set_error("NCCL error %d (%s) at %s:%d: %s", (int)_r, ncclGetErrorString(_r), __FILE__, __LINE__, #expr); \
            return 1000 + (int)_r;                                                            \
        }                                                                                     \
    } while (0)

// ------------------------------------------------------------------ peer memory (NVLink P2P) ----
namespace gcnk {
static thread_local const float *t_mirror_out = nullptr;
static thread_local Mirror t_mirror = {};

// Measured on 8 B200 (Reddit shape, 29 K rows per rank): the mirrored epilogues issue 4-16 byte remote stores per
// lane and made the feature transform 3x and the layer-2 kernel 3.4x slower than their 1/8 share, while one
// coalesced gcnk_peer_push of the finished rows takes ~10 us.  So the fused epilogues are opt-in
// (GCNK_MIRROR_EPILOGUE=1); by default a registration stays pending and the caller pushes.
Mirror take_mirror(const float *out) {
    static const bool enabled = getenv("GCNK_MIRROR_EPILOGUE") && atoi(getenv("GCNK_MIRROR_EPILOGUE")) != 0;
    Mirror m = {};
    if (enabled && out && t_mirror_out == out) { m = t_mirror; t_mirror_out = nullptr; t_mirror.n = 0; }
    return m;
}
}  // namespace gcnk

namespace gcnk {
// How long a kernel spins on a peer's flag before it raises *err and gives up (so that a lost peer becomes an error,
// not a hung GPU): GCN_PEER_TIMEOUT_S seconds (default 60; the first exchange of a run also absorbs whatever host-side
// skew the ranks have — uneven build(), first-use view construction, a paging stall).
long long peer_spin_cycles() {
    static const long long cycles = [] {
        const char *e = getenv("GCN_PEER_TIMEOUT_S");
        double sec = e && *e ? atof(e) : 60.0;
        if (!(sec > 0)) sec = 60.0;
        return (long long)(sec * 2.0e9);                    // clock64 ticks at <= 2 GHz
    }();
    return cycles;
}
}  // namespace gcnk

namespace {

struct FlagPtrs { int *p[8]; };

// One warp: lane r < world publishes `value` in peer r's flag slot for this rank, then waits until peer r has
// published it here.  Launched after a producer kernel with mirrored stores: when it completes, every rank's
// rows of the gather source are in this GPU's buffer.  Each rank runs on its own GPU, so the spin cannot
// starve the peer it waits for; a ~2 s timeout turns a lost peer into an error instead of a hang.
// the flag exchange of peer_barrier_kernel, run by the first `world` threads of the LAST CTA of a kernel to finish:
// every CTA fences its remote stores system-wide and then bumps *counter, so whatever the other CTAs wrote to the
// peers is visible there before the flag is
__device__ __forceinline__ bool last_cta_arrives(unsigned *counter) {
    __shared__ bool s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(counter, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) *counter = 0;                           // ready for the next launch
    }
    __syncthreads();
    return s_last;
}

__device__ __forceinline__ void flag_exchange(const FlagPtrs &flags, int rank, int world, int value, int *err, long long limit) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    volatile int *remote = flags.p[r] + rank;
    *remote = value;
    __threadfence_system();
    volatile int *mine = flags.p[rank] + r;
    const long long t0 = clock64();
    while (*mine < value) {
        if (clock64() - t0 > limit) { *err = 1; break; }
    }
    __threadfence_system();
}

// push + barrier in one launch
__global__ void __launch_bounds__(256) push_barrier_kernel(const float4 *__restrict__ src, Mirror m, size_t n_vec, FlagPtrs flags, int rank,
                                                           int world, int value, int *err, unsigned *counter, long long limit) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_vec; i += stride) {
        const float4 v = src[i];
        for (int p = 0; p < m.n; p++) reinterpret_cast<float4 *>(m.p[p])[i] = v;
    }
    if (last_cta_arrives(counter)) flag_exchange(flags, rank, world, value, err, limit);
}

// push + SIGNAL: the rows go to the peers and the last CTA to finish publishes `value` in each peer's flag slot for
// (this buffer, this rank) — nobody waits here.  The consumer (the next GraphSum gather on each rank) waits for the
// flags of the ranks it reads from at ITS start (gcnk_gather_wait_next), so a rank that is ahead runs on with its own
// work instead of idling in a barrier, and the barrier kernel and its launch disappear from the step.
// rows == nullptr: the contiguous block [0, n_vec) of float4; otherwise n_rows listed rows of vec_per_row float4 each,
// a separate list per peer (halo exchange: only the rows that peer's columns reference).
struct RowLists { const int *rows[MAX_PEERS]; int count[MAX_PEERS]; };
template <typename T>
__global__ void __launch_bounds__(256) push_signal_kernel(const T *__restrict__ src, Mirror m, size_t n_vec, FlagPtrs peer_flags, int value,
                                                          unsigned *counter) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_vec; i += stride) {
        const T v = src[i];
        for (int p = 0; p < m.n; p++) reinterpret_cast<T *>(m.p[p])[i] = v;
    }
    if (last_cta_arrives(counter) && (int)threadIdx.x < m.n) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(peer_flags.p[threadIdx.x]) = value;
    }
}
__global__ void __launch_bounds__(256) push_rows_signal_kernel(const float4 *__restrict__ src, Mirror m, RowLists lists, int vec_per_row,
                                                               FlagPtrs peer_flags, int value, unsigned *counter) {
    // one (row, float4) element per thread per step; peers take turns so that every list is walked coalesced
    for (int p = 0; p < m.n; p++) {
        const size_t total = (size_t)lists.count[p] * vec_per_row;
        const int *rows = lists.rows[p];
        float4 *dst = reinterpret_cast<float4 *>(m.p[p]);
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const size_t e = (size_t)rows[i / vec_per_row] * vec_per_row + i % vec_per_row;
            dst[e] = src[e];
        }
    }
    if (last_cta_arrives(counter) && (int)threadIdx.x < m.n) {
        __threadfence_system();
        *reinterpret_cast<volatile int *>(peer_flags.p[threadIdx.x]) = value;
    }
}

struct Segs { float *p[4]; unsigned count[4]; int n; };
struct Areas { float *p[8]; };

// all-reduce step 1: this rank's segments, packed, into slot[rank] of every rank's exchange area; then the barrier
__global__ void __launch_bounds__(256) allreduce_scatter_kernel(Segs segs, Areas areas, size_t slot_floats, FlagPtrs flags, int rank, int world,
                                                                int value, int *err, unsigned *counter, long long limit) {
    unsigned total = 0;
    for (int k = 0; k < segs.n; k++) total += segs.count[k];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        unsigned j = i;
        int k = 0;
        while (j >= segs.count[k]) { j -= segs.count[k]; k++; }
        const float v = segs.p[k][j];
        for (int r = 0; r < world; r++) areas.p[r][(size_t)rank * slot_floats + i] = v;
    }
    if (last_cta_arrives(counter)) flag_exchange(flags, rank, world, value, err, limit);
}

// all-reduce step 2: the slots summed in rank order (the same order on every rank => bit-identical results)
__global__ void __launch_bounds__(256) allreduce_gather_kernel(Segs segs, const float *__restrict__ area, size_t slot_floats, int world) {
    unsigned total = 0;
    for (int k = 0; k < segs.n; k++) total += segs.count[k];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < world; r++) s += area[(size_t)r * slot_floats + i];
        unsigned j = i;
        int k = 0;
        while (j >= segs.count[k]) { j -= segs.count[k]; k++; }
        segs.p[k][j] = s;
    }
}

__global__ void peer_barrier_kernel(FlagPtrs flags, int rank, int world, int value, int *err, long long limit) { flag_exchange(flags, rank, world, value, err, limit); }

__global__ void push_rows_kernel(const float4 *__restrict__ src, Mirror m, size_t n_vec) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_vec; i += stride) {
        const float4 v = src[i];
        for (int p = 0; p < m.n; p++) reinterpret_cast<float4 *>(m.p[p])[i] = v;
    }
}

}  // namespace

extern "C" {

int gcnk_ipc_export(const void *dptr, void *h_handle64) {
    GCNK_REQUIRE(dptr && h_handle64, "null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    GCNK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
    memcpy(h_handle64, &h, sizeof h);
    return GCNK_OK;
}

int gcnk_ipc_import(void **dptr, const void *h_handle64) {
    GCNK_REQUIRE(dptr && h_handle64, "null");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, sizeof h);
    GCNK_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GCNK_OK;
}

int gcnk_ipc_release(void *dptr) {
    if (dptr) GCNK_CUDA(cudaIpcCloseMemHandle(dptr));
    return GCNK_OK;
}

int gcnk_mirror_next(const float *local_out, float *const *peer_out, int n_peers) {
    GCNK_REQUIRE(local_out && (peer_out || n_peers == 0) && n_peers >= 0 && n_peers <= MAX_PEERS, "bad arguments");
    t_mirror_out = n_peers ? local_out : nullptr;
    t_mirror.n = n_peers;
    for (int i = 0; i < n_peers; i++) t_mirror.p[i] = peer_out[i];
    return GCNK_OK;
}

int gcnk_mirror_pending(const float *local_out) {
    const int pending = local_out && t_mirror_out == local_out && t_mirror.n > 0;
    if (pending) { t_mirror_out = nullptr; t_mirror.n = 0; }
    return pending;
}

int gcnk_peer_push(const float *local_rows, float *const *peer_rows, int n_peers, size_t n_floats, gcnk_stream_t stream) {
    GCNK_REQUIRE(local_rows && peer_rows && n_peers >= 0 && n_peers <= MAX_PEERS && n_floats % 4 == 0, "bad arguments");
    if (!n_peers || !n_floats) return GCNK_OK;
    Mirror m = {};
    m.n = n_peers;
    for (int i = 0; i < n_peers; i++) m.p[i] = peer_rows[i];
    const size_t n_vec = n_floats / 4;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((n_vec + 255) / 256, (size_t)sm_count() * 4));
    push_rows_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float4 *>(local_rows), m, n_vec);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_peer_push_barrier(const float *local_rows, float *const *peer_rows, int n_peers, size_t n_floats, int *const *flag_arrays,
                           int rank, int world, int value, int *d_err, unsigned *d_counter, gcnk_stream_t stream) {
    GCNK_REQUIRE(local_rows && peer_rows && n_peers >= 0 && n_peers <= MAX_PEERS && n_floats % 4 == 0 && flag_arrays && d_err && d_counter &&
                     world >= 1 && world <= 8,
                 "bad arguments");
    Mirror m = {};
    m.n = n_peers;
    for (int i = 0; i < n_peers; i++) m.p[i] = peer_rows[i];
    FlagPtrs f = {};
    for (int r = 0; r < world; r++) f.p[r] = flag_arrays[r];
    const size_t n_vec = n_floats / 4;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((n_vec + 255) / 256, (size_t)sm_count() * 2));
    push_barrier_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float4 *>(local_rows), m, n_vec, f, rank, world, value, d_err, d_counter, peer_spin_cycles());
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_peer_push_signal(const float *local_rows, float *const *peer_rows, int n_peers, size_t n_floats, const int *const *d_row_lists,
                          const int *h_row_counts, int dim, int *const *peer_flag_slots, int value, unsigned *d_counter, gcnk_stream_t stream) {
    GCNK_REQUIRE(local_rows && peer_rows && n_peers >= 0 && n_peers <= MAX_PEERS && peer_flag_slots && d_counter, "bad arguments");
    GCNK_REQUIRE(!d_row_lists || (h_row_counts && dim > 0 && dim % 4 == 0), "row lists need counts and a row width that is a multiple of 4");
    if (!n_peers) return GCNK_OK;
    Mirror m = {};
    FlagPtrs f = {};
    m.n = n_peers;
    for (int i = 0; i < n_peers; i++) { m.p[i] = peer_rows[i]; f.p[i] = peer_flag_slots[i]; }
    if (d_row_lists) {
        RowLists l = {};
        size_t most = 0;
        for (int i = 0; i < n_peers; i++) { l.rows[i] = d_row_lists[i]; l.count[i] = h_row_counts[i]; most = std::max(most, (size_t)h_row_counts[i] * (dim / 4)); }
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((most + 255) / 256, (size_t)sm_count() * 2));
        push_rows_signal_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float4 *>(local_rows), m, l, dim / 4, f, value, d_counter);
    } else {
        bool vec4 = n_floats % 4 == 0 && reinterpret_cast<uintptr_t>(local_rows) % 16 == 0;
        for (int i = 0; i < n_peers; i++) vec4 = vec4 && reinterpret_cast<uintptr_t>(peer_rows[i]) % 16 == 0;
        const size_t n_vec = vec4 ? n_floats / 4 : n_floats;      // unaligned blocks (a compact range of loss terms) go float by float
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((n_vec + 255) / 256, (size_t)sm_count() * 2));
        if (vec4) push_signal_kernel<float4><<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float4 *>(local_rows), m, n_vec, f, value, d_counter);
        else push_signal_kernel<float><<<grid, 256, 0, S(stream)>>>(local_rows, m, n_vec, f, value, d_counter);
    }
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_peer_allreduce(float *const *d_segs, const size_t *h_counts, int n_segs, float *const *slot_areas, size_t slot_floats,
                        int *const *flag_arrays, int rank, int world, int value, int *d_err, unsigned *d_counter, gcnk_stream_t stream) {
    GCNK_REQUIRE(d_segs && h_counts && n_segs > 0 && n_segs <= 4 && slot_areas && flag_arrays && d_err && d_counter && world >= 1 && world <= 8,
                 "bad arguments");
    Segs sg = {};
    sg.n = n_segs;
    size_t total = 0;
    for (int k = 0; k < n_segs; k++) { sg.p[k] = d_segs[k]; sg.count[k] = (unsigned)h_counts[k]; total += h_counts[k]; }
    GCNK_REQUIRE(total <= slot_floats, "segments do not fit the exchange slot");
    if (world == 1 || total == 0) return GCNK_OK;
    Areas ar = {};
    FlagPtrs f = {};
    for (int r = 0; r < world; r++) { ar.p[r] = slot_areas[r]; f.p[r] = flag_arrays[r]; }
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)sm_count()));
    allreduce_scatter_kernel<<<grid, 256, 0, S(stream)>>>(sg, ar, slot_floats, f, rank, world, value, d_err, d_counter, peer_spin_cycles());
    GCNK_LAUNCHED();
    allreduce_gather_kernel<<<grid, 256, 0, S(stream)>>>(sg, slot_areas[rank], slot_floats, world);
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_peer_barrier(int *const *flag_arrays, int rank, int world, int value, int *d_err, gcnk_stream_t stream) {
    GCNK_REQUIRE(flag_arrays && world >= 1 && world <= 8 && rank >= 0 && rank < world && d_err, "bad arguments");
    FlagPtrs f = {};
    for (int r = 0; r < world; r++) f.p[r] = flag_arrays[r];
    peer_barrier_kernel<<<1, 32, 0, S(stream)>>>(f, rank, world, value, d_err, peer_spin_cycles());
    GCNK_LAUNCHED();
    return GCNK_OK;
}

int gcnk_comm_allgather_bytes(gcnk_comm *c, const void *h_send, void *h_recv, int bytes_per_rank) {
    GCNK_REQUIRE(c && h_send && h_recv && bytes_per_rank > 0, "bad arguments");
    if (c->world == 1) { memcpy(h_recv, h_send, bytes_per_rank); return GCNK_OK; }
    char *d = nullptr;
    GCNK_CUDA(cudaMalloc(&d, (size_t)bytes_per_rank * c->world));
    GCNK_CUDA(cudaMemcpy(d + (size_t)c->rank * bytes_per_rank, h_send, bytes_per_rank, cudaMemcpyHostToDevice));
    GCNK_NCCL(ncclAllGather(d + (size_t)c->rank * bytes_per_rank, d, bytes_per_rank, ncclChar, c->comm, nullptr));
    GCNK_CUDA(cudaStreamSynchronize(nullptr));
    GCNK_CUDA(cudaMemcpy(h_recv, d, (size_t)bytes_per_rank * c->world, cudaMemcpyDeviceToHost));
    GCNK_CUDA(cudaFree(d));
    return GCNK_OK;
}

int gcnk_comm_unique_id(void *id128) {
    GCNK_REQUIRE(id128, "null");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    GCNK_NCCL(ncclGetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return GCNK_OK;
}

int gcnk_comm_create(gcnk_comm **out, const void *id128, int rank, int world, int device) {
    GCNK_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world, "bad arguments");
    GCNK_CUDA(cudaSetDevice(device));
    gcnk_comm *c = new gcnk_comm;
    c->rank = rank; c->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", ncclGetErrorString(r));
        delete c;
        return 1000 + (int)r;
    }
    *out = c;
    return GCNK_OK;
}

int gcnk_comm_destroy(gcnk_comm *c) {
    if (!c) return GCNK_OK;
    if (c->comm) ncclCommDestroy(c->comm);
    delete c;
    return GCNK_OK;
}

int gcnk_comm_rank(const gcnk_comm *c, int *rank, int *world) {
    GCNK_REQUIRE(c, "null");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return GCNK_OK;
}

int gcnk_comm_allgather_rows(gcnk_comm *c, float *d_all, const int *h_row_begin, int dim, gcnk_stream_t stream) {
    GCNK_REQUIRE(c && d_all && h_row_begin && dim > 0, "bad arguments");
    if (c->world == 1) return GCNK_OK;
    GCNK_NCCL(ncclGroupStart());
    for (int r = 0; r < c->world; r++) {
        const size_t off = (size_t)h_row_begin[r] * dim, cnt = (size_t)(h_row_begin[r + 1] - h_row_begin[r]) * dim;
        if (cnt) GCNK_NCCL(ncclBroadcast(d_all + off, d_all + off, cnt, ncclFloat, r, c->comm, S(stream)));
    }
    GCNK_NCCL(ncclGroupEnd());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GCNK_OK;
}

int gcnk_comm_allreduce(gcnk_comm *c, float *const *d_bufs, const size_t *counts, int n_bufs, int op_max, gcnk_stream_t stream) {
    GCNK_REQUIRE(c && d_bufs && counts && n_bufs > 0, "bad arguments");
    if (c->world == 1) return GCNK_OK;
    GCNK_NCCL(ncclGroupStart());
    for (int i = 0; i < n_bufs; i++)
        if (counts[i]) GCNK_NCCL(ncclAllReduce(d_bufs[i], d_bufs[i], counts[i], ncclFloat, op_max ? ncclMax : ncclSum, c->comm, S(stream)));
    GCNK_NCCL(ncclGroupEnd());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GCNK_OK;
}

}  // extern "C"
