#!/usr/bin/env python
"""tools/time_matmul.py — times gcnk_matmul_fw (CUDA events, after warm-up) at a products-like layer-2 shape."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cuda_gcn_b200 import abi  # noqa: E402

abi.require_device(0)
m, n, p = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (2449029 // 4, 256, 47)))
rng = np.random.default_rng(0)
a = abi.dev(rng.standard_normal((m, n)).astype(np.float32))
b = abi.dev((rng.standard_normal((n, p)) * 0.3).astype(np.float32))
c = abi.DeviceArray((m, p), np.float32)
for _ in range(3):
    abi.k.gcnk_matmul_fw(a.ptr, b.ptr, c.ptr, m, n, p, None)
e0, e1 = abi.Event(), abi.Event()
abi.k.gcnk_device_sync()
e0.record()
for _ in range(10):
    abi.k.gcnk_matmul_fw(a.ptr, b.ptr, c.ptr, m, n, p, None)
e1.record(); e1.sync()
ms = e0.elapsed_ms(e1) / 10
gb = (m * n + m * p) * 4 / 1e9
print(f"matmul_fw {m}x{n}x{p}: {ms * 1e3:.1f} us  {gb / ms * 1e3:.0f} GB/s  {2 * m * n * p / ms / 1e9:.2f} TFLOP/s (fp32-equivalent)  launches={abi.load().gcnk_launch_count()}")
