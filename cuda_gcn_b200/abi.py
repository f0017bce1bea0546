"""cuda_gcn_b200/abi.py — ctypes binding of libgcnk.so (include/gcnk.h), the C ABI of the CUDA kernels.

This is the thinnest possible Python face over the C ABI: it exists so that the parity tests and
bench.py can call exactly the entry points a C++/cgo/ctypes host would bind.  There is NO fallback:
a missing library raises ImportError-like RuntimeError at load time and every non-zero return code
raises GcnkError carrying gcnk_last_error().  Nothing here computes anything on the CPU.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libgcnk.so"

vp, i32, i64, f32, u64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t


class GcnkError(RuntimeError):
    pass


class CeResult(C.Structure):
    _fields_ = [("loss", C.c_float), ("count", C.c_int), ("wrong", C.c_int), ("pad", C.c_int)]


class AdamTensor(C.Structure):
    _fields_ = [("data", vp), ("grad", vp), ("m", vp), ("v", vp), ("size", C.c_int), ("decay", C.c_int)]


# name -> (restype, argtypes); every function include/gcnk.h declares is listed here
SIGNATURES = {
    "gcnk_version": (i32, []),
    "gcnk_last_error": (C.c_char_p, []),
    "gcnk_device_count": (i32, [C.POINTER(i32)]),
    "gcnk_set_device": (i32, [i32]),
    "gcnk_device_info": (i32, [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(sz), C.POINTER(i32)]),
    "gcnk_async_error": (i32, [vp]),
    "gcnk_async_error_flag": (i32, [C.POINTER(vp)]),
    "gcnk_launch_count": (i64, []),
    "gcnk_malloc": (i32, [C.POINTER(vp), sz]),
    "gcnk_free": (i32, [vp]),
    "gcnk_malloc_host": (i32, [C.POINTER(vp), sz]),
    "gcnk_free_host": (i32, [vp]),
    "gcnk_memcpy_h2d": (i32, [vp, vp, sz, vp]),
    "gcnk_memcpy_d2h": (i32, [vp, vp, sz, vp]),
    "gcnk_memcpy_d2d": (i32, [vp, vp, sz, vp]),
    "gcnk_memset": (i32, [vp, i32, sz, vp]),
    "gcnk_stream_create": (i32, [C.POINTER(vp)]),
    "gcnk_stream_create_low_priority": (i32, [C.POINTER(vp)]),
    "gcnk_stream_destroy": (i32, [vp]),
    "gcnk_stream_sync": (i32, [vp]),
    "gcnk_device_sync": (i32, []),
    "gcnk_event_create": (i32, [C.POINTER(vp)]),
    "gcnk_event_destroy": (i32, [vp]),
    "gcnk_event_record": (i32, [vp, vp]),
    "gcnk_event_sync": (i32, [vp]),
    "gcnk_event_elapsed_ms": (i32, [vp, vp, C.POINTER(f32)]),
    "gcnk_stream_wait_event": (i32, [vp, vp]),
    "gcnk_flush_l2": (i32, [vp]),
    "gcnk_graph_create": (i32, [C.POINTER(vp), vp, vp, i32, i64, i32, vp, vp]),
    "gcnk_graph_create_view": (i32, [C.POINTER(vp), vp, vp, vp, vp]),
    "gcnk_graph_destroy": (i32, [vp]),
    "gcnk_graph_dinv": (i32, [vp, C.POINTER(vp)]),
    "gcnk_graph_stats": (i32, [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "gcnk_graphsum": (i32, [vp, vp, vp, i32, vp]),
    "gcnk_graph_release_scratch": (i32, [vp]),
    "gcnk_mask_row_stride_bits": (i32, [i32]),
    "gcnk_gather_variant": (i32, [i32]),
    "gcnk_scale_rows": (i32, [vp, vp, vp, i32, i32, vp]),
    "gcnk_gather_plain": (i32, [vp, vp, vp, i32, vp]),
    "gcnk_gather_relu_drop": (i32, [vp, vp, vp, vp, vp, f32, i32, vp]),
    "gcnk_gather_mask": (i32, [vp, vp, vp, vp, f32, i32, vp]),
    "gcnk_spmat_create": (i32, [C.POINTER(vp), vp, vp, i32, i32, i64, vp]),
    "gcnk_spmat_destroy": (i32, [vp]),
    "gcnk_spmat_is_dense": (i32, [vp, C.POINTER(i32)]),
    "gcnk_spmm_fw": (i32, [vp, vp, vp, vp, i32, vp, f32, vp, vp]),
    "gcnk_spmm_bw": (i32, [vp, vp, vp, vp, i32, vp, f32, vp]),
    "gcnk_dense_transform": (i32, [vp, i32, i32, vp, vp, i32, vp, f32, vp, i32, vp]),
    "gcnk_dense_pack": (i32, [vp, i32, i32, vp, i32, vp]),
    "gcnk_dense_transform_ld": (i32, [vp, i32, i32, i32, vp, vp, i32, vp, f32, vp, i32, vp]),
    "gcnk_dense_transform_bw_workspace": (sz, [i32, i32]),
    "gcnk_dense_transform_bw_ld": (i32, [vp, i32, i32, i32, vp, vp, i32, vp, f32, vp, sz, vp]),
    "gcnk_matmul_fw": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "gcnk_matmul_bw_a": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "gcnk_matmul_bw_b": (i32, [vp, vp, vp, i32, i32, i32, vp, sz, vp]),
    "gcnk_matmul_bw_b_workspace": (sz, [i32, i32, i32]),
    "gcnk_relu_fw": (i32, [vp, vp, i64, i32, vp]),
    "gcnk_relu_bw": (i32, [vp, vp, i64, vp]),
    "gcnk_rng_create": (i32, [C.POINTER(vp), u64, u64]),
    "gcnk_rng_destroy": (i32, [vp]),
    "gcnk_rng_seed": (i32, [vp, C.c_long]),
    "gcnk_rng_get_state": (i32, [vp, C.POINTER(u64)]),
    "gcnk_rng_set_state": (i32, [vp, u64, u64]),
    "gcnk_rng_skip": (i32, [vp, u64]),
    "gcnk_rng_next_host": (i32, [vp, vp, i64]),
    "gcnk_dropout_mask": (i32, [vp, vp, i64, f32, vp]),
    "gcnk_dropout_apply": (i32, [vp, vp, i64, f32, vp]),
    "gcnk_softmax_ce": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, sz, vp]),
    "gcnk_softmax_ce_workspace": (sz, [i32, i32]),
    "gcnk_accuracy": (i32, [vp, vp, i32, i32, vp, vp]),
    "gcnk_set_truth": (i32, [vp, vp, vp, i32, i32, vp]),
    "gcnk_adam_step": (i32, [C.POINTER(AdamTensor), i32, f32, f32, f32, f32, f32, vp, vp]),
    "gcnk_sum_squares": (i32, [vp, i64, vp, vp]),
    "gcnk_layer2_fused": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "gcnk_layer2_fused_terms": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp, vp, vp]),
    "gcnk_dense_transform_tc": (i32, [vp, i32, i32, i32, vp, vp, i32, vp, f32, vp, i32, vp]),
    "gcnk_dense_transform_bw_tc_workspace": (sz, [i32, i32, i32]),
    "gcnk_dense_transform_bw_tc": (i32, [vp, i32, i32, i32, vp, vp, i32, vp, f32, vp, sz, vp]),
    "gcnk_matmul_nn": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, vp]),
    "gcnk_matmul_nt": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp]),
    "gcnk_matmul_tn_workspace": (sz, [i32, i32, i32]),
    "gcnk_matmul_tn": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, sz, vp]),
    "gcnk_drop_scale_rows": (i32, [vp, i32, i32, vp, f32, vp, vp, vp]),
    "gcnk_relu_dropout_fw": (i32, [vp, i64, vp, f32, vp, vp]),
    "gcnk_mask_scale_bw": (i32, [vp, i64, vp, f32, vp]),
    "gcnk_pad_cols": (i32, [vp, vp, i32, i32, i32, vp]),
    "gcnk_unpad_cols": (i32, [vp, vp, i32, i32, i32, vp]),
    "gcnk_ce_rows_workspace": (sz, [i32]),
    "gcnk_ce_rows": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp, vp, vp]),
    "gcnk_sequential_sum": (i32, [vp, i32, vp, f32, vp, i32, i32, i32, vp, vp]),
    "gcnk_layer2_workspace": (sz, [i32, i32, i32]),
    "gcnk_comm_unique_id": (i32, [vp]),
    "gcnk_comm_create": (i32, [C.POINTER(vp), vp, i32, i32, i32]),
    "gcnk_comm_destroy": (i32, [vp]),
    "gcnk_comm_rank": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
    "gcnk_comm_allgather_rows": (i32, [vp, vp, vp, i32, vp]),
    "gcnk_comm_allreduce": (i32, [vp, vp, vp, i32, i32, vp]),
    "gcnk_ipc_export": (i32, [vp, vp]),
    "gcnk_ipc_import": (i32, [C.POINTER(vp), vp]),
    "gcnk_ipc_release": (i32, [vp]),
    "gcnk_comm_allgather_bytes": (i32, [vp, vp, vp, i32]),
    "gcnk_mirror_next": (i32, [vp, vp, i32]),
    "gcnk_mirror_pending": (i32, [vp]),
    "gcnk_peer_push": (i32, [vp, vp, i32, sz, vp]),
    "gcnk_peer_barrier": (i32, [vp, i32, i32, i32, vp, vp]),
    "gcnk_peer_push_barrier": (i32, [vp, vp, i32, sz, vp, i32, i32, i32, vp, vp, vp]),
    "gcnk_peer_push_signal": (i32, [vp, vp, i32, sz, vp, vp, i32, vp, i32, vp, vp]),
    "gcnk_gather_wait_next": (i32, [vp, i32, i32, i32, vp]),
    "gcnk_graph_rotate": (i32, [vp, i32, i32, vp]),
    "gcnk_gather_exchange_next": (i32, [vp, sz, vp, i32, vp, vp, i32, i32, i32, vp, vp]),
    "gcnk_gather_raw": (i32, [vp, vp, vp, i32, vp]),
    "gcnk_gather_init_next": (i32, [vp]),
    "gcnk_peer_allreduce": (i32, [vp, vp, i32, vp, sz, vp, i32, i32, i32, vp, vp, vp]),
    "gcnk_partition_rows": (i32, [vp, i32, i32, vp]),
}

# functions whose return value is not an error code
_NOT_RC = {"gcnk_dense_transform_bw_workspace", "gcnk_mirror_pending", "gcnk_version", "gcnk_last_error", "gcnk_launch_count", "gcnk_matmul_bw_b_workspace",
           "gcnk_softmax_ce_workspace", "gcnk_layer2_workspace", "gcnk_mask_row_stride_bits", "gcnk_gather_variant", "gcnk_matmul_tn_workspace",
           "gcnk_ce_rows_workspace", "gcnk_async_error", "gcnk_dense_transform_bw_tc_workspace"}

_lib = None


def load():
    """dlopen libgcnk.so and bind every symbol.  Raises if the library is not built — loudly, by design."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise GcnkError(f"{LIB_PATH} is not built (run `make -C cuda_gcn_b200/csrc` or __graft_entry__.build()); "
                        "there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)           # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise GcnkError(f"{what}: rc={rc}: {load().gcnk_last_error().decode(errors='replace')}")


class _Checked:
    """Attribute access returns the C function wrapped so that a non-zero return code raises."""

    def __getattr__(self, name):
        fn = getattr(load(), name)
        if name in _NOT_RC:
            return fn

        def call(*a):
            check(fn(*a), name)
        call.__name__ = name
        setattr(self, name, call)
        return call


k = _Checked()


def device_count() -> int:
    n = i32(0)
    rc = load().gcnk_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def require_device(index: int = 0):
    if device_count() <= index:
        raise GcnkError("no CUDA device visible: the engine has no CPU path")
    k.gcnk_set_device(index)


class DeviceArray:
    """A device allocation with a numpy dtype/shape; copies are synchronous on the given stream."""

    def __init__(self, shape, dtype, stream=None):
        self.shape = tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, tuple) else shape
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = vp()
        k.gcnk_malloc(C.byref(p), self.nbytes)
        self.ptr = p.value
        self.stream = stream

    @classmethod
    def from_numpy(cls, a, stream=None):
        a = np.ascontiguousarray(a)
        d = cls(a.shape, a.dtype, stream)
        d.upload(a)
        return d

    @classmethod
    def zeros(cls, shape, dtype, stream=None):
        d = cls(shape, dtype, stream)
        k.gcnk_memset(d.ptr, 0, d.nbytes, stream)
        return d

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.nbytes, self.nbytes)
        k.gcnk_memcpy_h2d(self.ptr, a.ctypes.data, self.nbytes, self.stream)
        k.gcnk_stream_sync(self.stream)

    def numpy(self):
        out = np.empty(self.shape, self.dtype)
        k.gcnk_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, self.stream)
        k.gcnk_stream_sync(self.stream)
        return out

    def free(self):
        if self.ptr:
            k.gcnk_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def dev(a, dtype=None):
    """numpy -> DeviceArray"""
    return DeviceArray.from_numpy(np.asarray(a, dtype=dtype) if dtype is not None else a)


class Graph:
    """gcnk_graph handle over device CSR arrays (kept alive here)."""

    def __init__(self, indptr, indices, n_cols=None, dinv_global=None):
        self.indptr = dev(indptr, np.int32)
        # four entries of padding: the int4 index reads of the gather round the array length up (gcnk_gather_variant)
        self.indices = dev(np.concatenate([np.asarray(indices, np.int32), np.full(4, -1, np.int32)]), np.int32)
        self.n = len(indptr) - 1
        self.nnz = int(len(indices))
        self.n_cols = self.n if n_cols is None else n_cols
        self.dinv_global = dev(dinv_global, np.float32) if dinv_global is not None else None
        h = vp()
        k.gcnk_graph_create(C.byref(h), self.indptr.ptr, self.indices.ptr, self.n, self.nnz, self.n_cols,
                            self.dinv_global.ptr if self.dinv_global is not None else None, None)
        self.h = h.value

    def dinv_ptr(self):
        p = vp()
        k.gcnk_graph_dinv(self.h, C.byref(p))
        return p.value

    def dinv(self):
        out = np.empty(self.n, np.float32)
        k.gcnk_memcpy_d2h(out.ctypes.data, self.dinv_ptr(), out.nbytes, None)
        k.gcnk_stream_sync(None)
        return out

    def stats(self):
        n, nnz, md, sym, nb = i32(), i64(), i32(), i32(), i32()
        k.gcnk_graph_stats(self.h, C.byref(n), C.byref(nnz), C.byref(md), C.byref(sym), C.byref(nb))
        return dict(n=n.value, nnz=nnz.value, max_degree=md.value, symmetric=bool(sym.value), n_bins=nb.value)

    def close(self):
        if self.h:
            k.gcnk_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpMat:
    def __init__(self, indptr, indices, m, n):
        self.indptr = dev(indptr, np.int32)
        # four entries of padding: the int4 index reads of the gather round the array length up (gcnk_gather_variant)
        self.indices = dev(np.concatenate([np.asarray(indices, np.int32), np.full(4, -1, np.int32)]), np.int32)
        self.m, self.n, self.nnz = m, n, int(len(indices))
        h = vp()
        k.gcnk_spmat_create(C.byref(h), self.indptr.ptr, self.indices.ptr, m, n, self.nnz, None)
        self.h = h.value

    def is_dense(self):
        v = i32()
        k.gcnk_spmat_is_dense(self.h, C.byref(v))
        return bool(v.value)

    def close(self):
        if self.h:
            k.gcnk_spmat_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Rng:
    def __init__(self, s0=1, s1=2):
        h = vp()
        k.gcnk_rng_create(C.byref(h), s0, s1)
        self.h = h.value

    def seed(self, seed):
        k.gcnk_rng_seed(self.h, seed)

    def state(self):
        out = (u64 * 2)()
        k.gcnk_rng_get_state(self.h, out)
        return int(out[0]), int(out[1])

    def set_state(self, a, b):
        k.gcnk_rng_set_state(self.h, a, b)

    def skip(self, n):
        k.gcnk_rng_skip(self.h, n)

    def next_host(self, n):
        out = np.empty(n, np.uint32)
        k.gcnk_rng_next_host(self.h, out.ctypes.data, n)
        return out

    def close(self):
        if self.h:
            k.gcnk_rng_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Event:
    def __init__(self):
        h = vp()
        k.gcnk_event_create(C.byref(h))
        self.h = h.value

    def record(self, stream=None):
        k.gcnk_event_record(self.h, stream)

    def sync(self):
        k.gcnk_event_sync(self.h)

    def elapsed_ms(self, later: "Event") -> float:
        ms = f32()
        k.gcnk_event_elapsed_ms(self.h, later.h, C.byref(ms))
        return ms.value
