"""world_size-2 (and 3) gloo run on CPU of the host side of the row-partitioned path (tools/dist_cpu_check.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_rendezvous_over_gloo(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), str(ROOT / "tools" / "dist_cpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "DIST_CPU_OK" in out.stdout
