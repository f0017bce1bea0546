#!/usr/bin/env python
"""tools/time_gather_slice.py — one rank's share of the Reddit-shape GraphSum on ONE GPU: the row slice a rank of a
P-way partition owns (global column ids, full [N x 16] source), timed with CUDA events.  Separates what the slice
costs by itself from what the multi-GPU exchange adds."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cuda_gcn_b200 import abi, host_api  # noqa: E402

parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
abi.require_device(0)
d = host_api.Data.synth("reddit", 1.0)
a = d.arrays()
indptr, indices = a["graph_indptr"], a["graph_indices"]
n = len(indptr) - 1
dinv = (np.float32(1) / np.sqrt(np.diff(indptr).astype(np.float32))).astype(np.float32)
cuts = np.zeros(parts + 1, np.int32)
abi.k.gcnk_partition_rows(indptr.ctypes.data, n, parts, cuts.ctypes.data)
x = abi.dev(np.random.default_rng(0).standard_normal((n, 16)).astype(np.float32))
for r in (0, parts // 2):
    lo, hi = int(cuts[r]), int(cuts[r + 1])
    g = abi.Graph((indptr[lo:hi + 1] - indptr[lo]).astype(np.int32), indices[indptr[lo]:indptr[hi]], n_cols=n, dinv_global=dinv)
    out = abi.DeviceArray((hi - lo, 16), np.float32)
    for _ in range(3):
        abi.k.gcnk_gather_plain(g.h, x.ptr, out.ptr, 16, None)
    e0, e1 = abi.Event(), abi.Event()
    abi.k.gcnk_device_sync()
    e0.record()
    for _ in range(20):
        abi.k.gcnk_gather_plain(g.h, x.ptr, out.ptr, 16, None)
    e1.record(); e1.sync()
    us = e0.elapsed_ms(e1) / 20 * 1e3
    nnz = int(indptr[hi] - indptr[lo])
    print(f"rank {r}/{parts}: rows {hi - lo}, nnz {nnz}, gather {us:.1f} us = {nnz / us / 1e3:.0f} G edges/s  stats {g.stats()}")
