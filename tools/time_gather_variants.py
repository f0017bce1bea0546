#!/usr/bin/env python
"""tools/time_gather_variants.py — the full-graph GraphSum gather (dim 16) at Reddit shape under each index-fetch
variant of gcnk_gather_variant(): CUDA-event time per launch (L2 flushed between launches or not), the largest
difference against variant 0, and a float64 check of sampled rows."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cuda_gcn_b200 import abi, host_api  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
abi.require_device(0)
d = host_api.Data.synth("reddit", scale)
a = d.arrays()
indptr, indices = a["graph_indptr"], a["graph_indices"]
n, dim = len(indptr) - 1, 16
x = np.random.default_rng(0).standard_normal((n, dim)).astype(np.float32)
dx, out = abi.dev(x), abi.DeviceArray((n, dim), np.float32)
rows = np.concatenate([np.random.default_rng(1).integers(0, n, 64), [int(np.argmax(np.diff(indptr)))]])
dinv = 1.0 / np.sqrt(np.diff(indptr).astype(np.float64))
ref = np.stack([dinv[i] * x[indices[indptr[i]:indptr[i + 1]]].astype(np.float64).sum(0) for i in rows])
base = None
combos = [(0, None), (1, None), (2, None)]
if len(sys.argv) > 3:      # e.g. "0:256,0:320,2:240,2:288": variant:bins-per-SM of the static schedule
    combos = [(int(c.split(":")[0]), int(c.split(":")[1])) for c in sys.argv[3].split(",")]
for v, bins in combos:
    abi.k.gcnk_gather_variant(v)
    if bins is None:
        os.environ.pop("GCNK_GATHER_BINS_PER_SM", None)
    else:
        os.environ["GCNK_GATHER_BINS_PER_SM"] = str(bins)
    g = abi.Graph(indptr, indices)                       # the schedule is built here
    for _ in range(3):
        abi.k.gcnk_gather_plain(g.h, dx.ptr, out.ptr, dim, None)
    e0, e1 = abi.Event(), abi.Event()
    abi.k.gcnk_device_sync()
    e0.record()
    for _ in range(reps):
        abi.k.gcnk_gather_plain(g.h, dx.ptr, out.ptr, dim, None)
    e1.record(); e1.sync()
    us = e0.elapsed_ms(e1) / reps * 1e3
    got = out.numpy().reshape(n, dim)
    err = np.abs(got[rows] - ref).max() / np.abs(ref).max()
    if base is None:
        base = got.copy()
    print(f"variant {v} bins/SM {bins or 'default':>7}: {us:8.1f} us/launch  {len(indices) / us / 1e3:6.1f} G edges/s   max|diff vs first| "
          f"{np.abs(got - base).max():.3e}   sampled rows vs float64: {err:.2e}   {g.stats()}", flush=True)
    del g
abi.k.gcnk_gather_variant(0)
