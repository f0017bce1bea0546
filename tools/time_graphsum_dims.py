#!/usr/bin/env python
"""tools/time_graphsum_dims.py — GraphSum (gcnk_graphsum = pre-scale + gather, what CUDAGraphSum::forward/backward do,
cuda_module.cu:74-101) per call at each width the reference's models use (metric 2): CUDA-event time,
algorithmic GB/s (B_min = 4 nnz + 4 (N+1) + 8 N dim) and its fraction of the measured HBM peak.

    python tools/time_graphsum_dims.py [preset=reddit] [scale=1.0] [dims=16,41,47,256] [reps=10]

Forward and backward are the same launch on a symmetric graph (module.cpp:103-119 reuses the forward loop)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda_gcn_b200 import abi, host_api  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else "reddit"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dims = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "16,41,47,256").split(",")]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
peak = 6539.9
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass

abi.require_device(0)
data = host_api.Data.synth(preset, scale)                  # keep alive: arrays() are views
a = data.arrays()
indptr, indices = a["graph_indptr"], a["graph_indices"]
n, nnz = len(indptr) - 1, len(indices)
g = abi.Graph(indptr, indices)
print(f"{preset} x{scale}: n {n} nnz {nnz}  {g.stats()}  HBM peak {peak:.0f} GB/s")
for dim in dims:
    x = abi.dev(np.random.default_rng(dim).standard_normal((n, dim)).astype(np.float32))
    out = abi.DeviceArray((n, dim), np.float32)
    for _ in range(3):
        abi.k.gcnk_graphsum(g.h, x.ptr, out.ptr, dim, None)
    e0, e1 = abi.Event(), abi.Event()
    abi.k.gcnk_device_sync()
    e0.record()
    for _ in range(reps):
        abi.k.gcnk_graphsum(g.h, x.ptr, out.ptr, dim, None)
    e1.record(); e1.sync()
    us = e0.elapsed_ms(e1) / reps * 1e3
    b_min = 4 * nnz + 4 * (n + 1) + 8 * n * dim
    b_call = b_min + 8 * n * dim                            # + the pre-scale pass of gcnk_graphsum (read + write of [n x dim])
    print(f"dim {dim:4d}: {us:9.1f} us/call   B_min {b_min / 1e6:8.1f} MB -> {b_min / us / 1e3:7.1f} GB/s = {b_min / us / 1e3 / peak:.3f} of peak"
          f"   (incl. pre-scale pass: {b_call / 1e6:.1f} MB)", flush=True)
    del x, out
