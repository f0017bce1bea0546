"""cuda_gcn_b200/host_api.py — ctypes binding of libgcnhost.so (include/gcn_host.h): the C face of the
C++ host layer (Parser, GCNData, GCN training loop, timers).  Used by the tests and bench.py only;
it computes nothing itself and has no fallback (a missing library or GPU is an error)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libgcnhost.so"


class Params(C.Structure):
    _fields_ = [("num_nodes", C.c_int), ("input_dim", C.c_int), ("hidden_dim", C.c_int), ("output_dim", C.c_int),
                ("dropout", C.c_float), ("learning_rate", C.c_float), ("weight_decay", C.c_float),
                ("epochs", C.c_int), ("early_stopping", C.c_int)]


vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")

SIGNATURES = {
    "gcnh_default_params": (Params, []),
    "gcnh_data_new": (vp, []),
    "gcnh_data_free": (None, [vp]),
    "gcnh_data_parse": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.POINTER(Params), C.c_int]),
    "gcnh_data_fill": (C.c_int, [vp, C.c_int, _i32, _i32, _i32, _i32, _f32, _i32, _i32]),
    "gcnh_data_synth": (C.c_int, [vp, C.c_char_p, C.c_double, C.c_uint64, C.POINTER(Params)]),
    "gcnh_data_slice": (vp, [vp, C.c_int, C.c_int, ip, ip]),
    "gcnh_data_sizes": (None, [vp, C.POINTER(C.c_int64)]),
    "gcnh_data_graph_indptr": (ip, [vp]),
    "gcnh_data_graph_indices": (ip, [vp]),
    "gcnh_data_feature_indptr": (ip, [vp]),
    "gcnh_data_feature_indices": (ip, [vp]),
    "gcnh_data_feature_value": (fp, [vp]),
    "gcnh_data_label": (ip, [vp]),
    "gcnh_data_split": (ip, [vp]),
    "gcnh_engine_create": (vp, [C.POINTER(Params), vp, C.c_long, C.c_int, C.c_int]),
    "gcnh_comm_unique_id": (C.c_int, [vp]),
    "gcnh_engine_create_dist": (vp, [C.POINTER(Params), vp, C.c_long, C.c_int, C.c_int, C.c_int, vp]),
    "gcnh_engine_allreduce_host": (None, [vp, vp, C.c_int, C.c_int]),
    "gcnh_engine_destroy": (None, [vp]),
    "gcnh_engine_plan": (C.c_int, [vp]),
    "gcnh_engine_train_epoch": (None, [vp, fp, fp]),
    "gcnh_engine_eval": (None, [vp, C.c_int, fp, fp]),
    "gcnh_engine_epoch": (None, [vp, C.c_int, fp, fp, fp, fp]),
    "gcnh_engine_last_counts": (None, [vp, ip, ip]),
    "gcnh_engine_run": (C.c_int, [vp, C.c_int]),
    "gcnh_engine_set_input_host": (None, [vp, vp]),
    "gcnh_engine_epoch_prefetch": (None, [vp, C.c_int, vp, fp, fp, fp, fp]),
    "gcnh_engine_var_size": (C.c_int64, [vp, C.c_int]),
    "gcnh_engine_get_var": (None, [vp, C.c_int, C.c_int, _f32]),
    "gcnh_timer_enable_gpu": (None, [C.c_int]),
    "gcnh_timer_enable_mask": (None, [C.c_uint]),
    "gcnh_timer_slot": (C.c_int, [C.c_char_p]),
    "gcnh_timer_reset": (None, []),
    "gcnh_timer_total": (C.c_float, [C.c_int]),
    "gcnh_timer_calls": (C.c_int, [C.c_int]),
    "gcnh_timer_name": (C.c_char_p, [C.c_int]),
    "gcnh_timer_count": (C.c_int, []),
    "gcnh_alloc_pinned": (vp, [C.c_int64]),
    "gcnh_free_pinned": (None, [vp]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is not built (make -C cuda_gcn_b200/host); there is no fallback")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


PLAN_AUTO, PLAN_MODULES, PLAN_FUSED = 0, 1, 2


class Data:
    """GCNData (gcn.h:16-22) living in the C++ host layer; numpy views are zero-copy and read-only by convention."""

    def __init__(self):
        self.L = load()
        self.h = self.L.gcnh_data_new()
        self.params = self.L.gcnh_default_params()

    @classmethod
    def parse(cls, root, name, quiet=True):
        d = cls()
        ok = d.L.gcnh_data_parse(d.h, str(root).encode() if root else None, name.encode(), C.byref(d.params), int(quiet))
        return d if ok else None

    @classmethod
    def from_arrays(cls, gd):
        """gd: oracle.checker.GraphData-like (numpy arrays)."""
        d = cls()
        c = lambda a, t: np.ascontiguousarray(a, dtype=t)
        d.L.gcnh_data_fill(d.h, gd.num_nodes, c(gd.graph_indptr, np.int32), c(gd.graph_indices, np.int32),
                           c(gd.feature_indptr, np.int32), c(gd.feature_indices, np.int32), c(gd.feature_value, np.float32),
                           c(gd.label, np.int32), c(gd.split, np.int32))
        d.params.num_nodes, d.params.input_dim, d.params.output_dim = gd.num_nodes, gd.input_dim, gd.output_dim
        return d

    @classmethod
    def synth(cls, preset, scale=1.0, seed=0):
        d = cls()
        if not d.L.gcnh_data_synth(d.h, preset.encode(), float(scale), int(seed), C.byref(d.params)):
            raise RuntimeError(f"synthetic preset {preset!r} failed")
        return d

    def slice(self, rank, world):
        """(Data holding this rank's rows, row_begin, row_end) under the nnz-balanced row partition."""
        a, b = C.c_int(), C.c_int()
        out = Data.__new__(Data)
        out.L = self.L
        out.h = self.L.gcnh_data_slice(self.h, rank, world, C.byref(a), C.byref(b))
        if not out.h:
            raise RuntimeError("gcnh_data_slice failed")
        out.params = Params.from_buffer_copy(bytes(self.params))
        out.params.num_nodes = b.value - a.value
        return out, a.value, b.value

    def sizes(self):
        s = (C.c_int64 * 7)()
        self.L.gcnh_data_sizes(self.h, s)
        keys = ("num_nodes", "graph_nnz", "feature_nnz", "n_label", "n_split", "max_degree", "feature_rows")
        return dict(zip(keys, [int(v) for v in s]))

    def arrays(self):
        """Zero-copy numpy views of the GCNData vectors.  Each view keeps this object alive (numpy array -> ctypes
        array -> Data), so `Data.synth(...).arrays()` is safe; close() while views are in use is not."""
        s = self.sizes()
        owner = self

        def view(ptr, n, dt):
            if not n:
                return np.zeros(0, dt)
            ctype = C.c_float if dt == np.float32 else C.c_int32
            buf = (ctype * n).from_address(C.cast(ptr, C.c_void_p).value)
            buf._owner = owner                                # the ctypes array is the numpy view's base
            return np.frombuffer(buf, dtype=dt)
        L, h = self.L, self.h
        return dict(graph_indptr=view(L.gcnh_data_graph_indptr(h), s["num_nodes"] + 1, np.int32),
                    graph_indices=view(L.gcnh_data_graph_indices(h), s["graph_nnz"], np.int32),
                    feature_indptr=view(L.gcnh_data_feature_indptr(h), s["feature_rows"] + 1, np.int32),
                    feature_indices=view(L.gcnh_data_feature_indices(h), s["feature_nnz"], np.int32),
                    feature_value=view(L.gcnh_data_feature_value(h), s["feature_nnz"], np.float32),
                    label=view(L.gcnh_data_label(h), s["n_label"], np.int32),
                    split=view(L.gcnh_data_split(h), s["n_split"], np.int32))

    def close(self):
        if self.h:
            self.L.gcnh_data_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """GCN (gcn.h:24-44) on one GPU."""

    def __init__(self, data: Data, hidden_dim=16, dropout=0.5, lr=0.01, weight_decay=5e-4, epochs=100, early_stopping=0,
                 seed=1, plan=PLAN_AUTO, device=0, rank=0, world=1, nccl_id=None):
        """world > 1: row-partitioned engine; every rank passes the same data and seed, nccl_id = the 128 bytes
        rank 0 got from unique_id() (see rendezvous())."""
        self.L, self.data = load(), data
        p = Params(data.params.num_nodes, data.params.input_dim, hidden_dim, data.params.output_dim, dropout, lr,
                   weight_decay, epochs, early_stopping)
        self.params = p
        if world > 1:
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            self.h = self.L.gcnh_engine_create_dist(C.byref(p), data.h, seed, device, rank, world, buf)
        else:
            self.h = self.L.gcnh_engine_create(C.byref(p), data.h, seed, plan, device)

    def allreduce_host(self, values, op_max=False):
        a = np.ascontiguousarray(values, dtype=np.float32)
        self.L.gcnh_engine_allreduce_host(self.h, a.ctypes.data, len(a), int(op_max))
        return a

    @property
    def plan(self):
        return self.L.gcnh_engine_plan(self.h)

    def train_epoch(self):
        a, b = C.c_float(), C.c_float()
        self.L.gcnh_engine_train_epoch(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def eval(self, split):
        a, b = C.c_float(), C.c_float()
        self.L.gcnh_engine_eval(self.h, split, C.byref(a), C.byref(b))
        return a.value, b.value

    def epoch(self, eval_split=2):
        """train_epoch() + eval(eval_split) with one host sync; returns (train_loss, train_acc, eval_loss, eval_acc)."""
        v = [C.c_float() for _ in range(4)]
        self.L.gcnh_engine_epoch(self.h, eval_split, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def last_counts(self):
        a, b = C.c_int(), C.c_int()
        self.L.gcnh_engine_last_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def run(self):
        return self.L.gcnh_engine_run(self.h, 0)

    def set_input_host(self, ptr):
        self.L.gcnh_engine_set_input_host(self.h, ptr)

    def epoch_prefetch(self, eval_split, next_ptr):
        """epoch() on the current input while the pinned buffer at next_ptr is uploaded for the following pass."""
        v = [C.c_float() for _ in range(4)]
        self.L.gcnh_engine_epoch_prefetch(self.h, eval_split, next_ptr, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def var(self, idx, grad=False):
        out = np.zeros(self.L.gcnh_engine_var_size(self.h, idx), np.float32)
        if len(out):
            self.L.gcnh_engine_get_var(self.h, idx, int(grad), out)
        return out

    def close(self):
        if self.h:
            self.L.gcnh_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    if not load().gcnh_comm_unique_id(buf):
        raise RuntimeError("ncclGetUniqueId failed")
    return buf.raw


def rendezvous(rank: int, world: int, addr: str = "127.0.0.1", port: int = 29500, timeout: float = 120.0, uid: bytes | None = None) -> bytes:
    """Share rank 0's ncclUniqueId with the other ranks of ONE node over a loopback TCP socket (stdlib only).
    Rank 0 listens on the first free port of port+1..port+32; the others probe the same sequence.  A 6-byte
    magic guards against unrelated services on those ports."""
    import socket
    import time
    magic = b"GCNKID"
    if world <= 1:
        return b""
    if rank == 0:
        uid = uid if uid is not None else unique_id()
        srv = None
        for off in range(1, 33):
            try:
                srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                srv.bind((addr, port + off))
                break
            except OSError:
                srv.close()
                srv = None
        if srv is None:
            raise RuntimeError("rendezvous: no free port")
        srv.listen(world)
        srv.settimeout(timeout)
        served = 0
        while served < world - 1:
            conn, _ = srv.accept()
            conn.settimeout(10)
            try:
                if conn.recv(6) == magic:
                    conn.sendall(magic + uid)
                    served += 1
            finally:
                conn.close()
        srv.close()
        return uid
    deadline = time.time() + timeout
    while time.time() < deadline:
        for off in range(1, 33):
            try:
                with socket.create_connection((addr, port + off), timeout=2) as c:
                    c.sendall(magic)
                    data = b""
                    while len(data) < 6 + 128:
                        chunk = c.recv(6 + 128 - len(data))
                        if not chunk:
                            break
                        data += chunk
                    if len(data) == 6 + 128 and data[:6] == magic:
                        return data[6:]
            except OSError:
                continue
        time.sleep(0.2)
    raise RuntimeError("rendezvous: rank 0 did not answer")


def timers():
    L = load()
    out = {}
    for t in range(L.gcnh_timer_count()):
        calls = L.gcnh_timer_calls(t)
        if calls:
            out[L.gcnh_timer_name(t).decode()] = (L.gcnh_timer_total(t), calls)
    return out
