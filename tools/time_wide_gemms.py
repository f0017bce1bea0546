#!/usr/bin/env python
"""tools/time_wide_gemms.py — the five dense products of the wide plan (host/gcn_wide.cpp) at ogbn-products widths, per
call (CUDA events after warm-up): algorithmic bytes / time against the measured HBM peak, and fp32-equivalent TFLOP/s.

    python tools/time_wide_gemms.py [rows=612257] [reps=10]
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda_gcn_b200 import abi  # noqa: E402

abi.require_device(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 2449029 // 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
peak = 6539.9
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
rng = np.random.default_rng(0)
F, H, C = 100, 256, 48
K = abi.k


def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = abi.Event(), abi.Event()
    K.gcnk_device_sync()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.sync()
    return e0.elapsed_ms(e1) / reps * 1e3


ax = abi.dev(rng.standard_normal((m, F)).astype(np.float32))
w1 = abi.dev((rng.standard_normal((F, H)) * 0.1).astype(np.float32))
h1 = abi.DeviceArray((m, H), np.float32)
w2 = abi.dev((rng.standard_normal((H, C)) * 0.1).astype(np.float32))
t = abi.DeviceArray((m, C), np.float32)
dinv = abi.dev((rng.random(m) + 0.5).astype(np.float32))
dt = abi.dev(rng.standard_normal((m, C)).astype(np.float32))
dh = abi.DeviceArray((m, H), np.float32)
dw1, dw2 = abi.DeviceArray((F, H), np.float32), abi.DeviceArray((H, C), np.float32)
wsb = max(K.gcnk_matmul_tn_workspace(m, F, H), K.gcnk_matmul_tn_workspace(m, H, C))
ws = abi.DeviceArray((max(wsb, 16) // 4,), np.float32)
cases = [
    ("nn  Z1 = AXd W1      [m x 100] x [100 x 256]", lambda: K.gcnk_matmul_nn(ax.ptr, F, w1.ptr, H, h1.ptr, H, m, F, H, None, None), m * (F + H) * 4, 2 * m * F * H),
    ("nn  T  = H1 W2 (.)d  [m x 256] x [256 x 48] ", lambda: K.gcnk_matmul_nn(h1.ptr, H, w2.ptr, C, t.ptr, C, m, H, C, dinv.ptr, None), m * (H + C) * 4, 2 * m * H * C),
    ("nt  dH1 = dT W2^T    [m x 48] x [256 x 48]^T", lambda: K.gcnk_matmul_nt(dt.ptr, C, w2.ptr, C, dh.ptr, H, m, C, H, None), m * (C + H) * 4, 2 * m * H * C),
    ("tn  dW2 = H1^T dT    [m x 256]^T x [m x 48] ", lambda: K.gcnk_matmul_tn(h1.ptr, H, dt.ptr, C, dw2.ptr, C, m, H, C, ws.ptr, wsb, None), m * (H + C) * 4, 2 * m * H * C),
    ("tn  dW1 = AXd^T dZ1  [m x 100]^T x [m x 256]", lambda: K.gcnk_matmul_tn(ax.ptr, F, dh.ptr, H, dw1.ptr, H, m, F, H, ws.ptr, wsb, None), m * (F + H) * 4, 2 * m * F * H),
]
print(f"rows {m}, HBM peak {peak:.0f} GB/s")
for name, fn, nbytes, flops in cases:
    us = timed(fn)
    print(f"{name}: {us:9.1f} us  {nbytes / us / 1e3:7.0f} GB/s = {nbytes / us / 1e3 / peak:.2f} of HBM peak  {flops / us / 1e6:6.1f} TFLOP/s fp32-equivalent", flush=True)
assert K.gcnk_async_error(None) == 0
