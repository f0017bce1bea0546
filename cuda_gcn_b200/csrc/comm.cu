// comm.cu — the exchange steps of the row-partitioned engine over NCCL (NVLink 5 / NVSwitch).
// The reference has no multi-GPU path at all; these entry points are what the partitioned
// plan adds: an all-gather of the GraphSum input with per-rank row counts (the partition is
// nnz-balanced, not row-balanced), and sum / max all-reduces for weight gradients and scalars.
//
// The all-gather is IN PLACE: every rank's producing kernel writes its rows straight into its slice of
// the [N x dim] buffer, and one grouped set of ncclBroadcast calls (one root per rank) fills the rest,
// so there is no staging copy.  The rendezvous (sharing the 128-byte ncclUniqueId) is the caller's job.
#include <nccl.h>

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

extern "C" {

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
