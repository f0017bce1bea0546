"""Row-partitioned engine == single-GPU engine ("2/4/8-rank run == 1-rank run, loss to 1e-6 rel.,
integer outputs identical").  Needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def n_gpus():
    from cuda_gcn_b200 import abi
    return abi.device_count()


@pytest.mark.parametrize("preset,scale,dropout,hidden", [("pubmed", 1.0, 0.5, 16), ("cora", 1.0, 0.0, 16), ("reddit", 0.02, 0.5, 16),
                                                         ("products", 0.004, 0.5, 256)])    # hidden 256: the wide plan
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_equals_single(preset, scale, dropout, hidden, world):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), str(ROOT / "tools" / "dist_check.py"), "--preset", preset, "--scale", str(scale),
           "--epochs", "6", "--dropout", str(dropout), "--hidden", str(hidden)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["replicated"], "ranks hold different weights"
    for d, s in zip(r["dist"], r["single"]):
        assert abs(d[0] - s[0]) <= 2e-6 * abs(s[0]) + 1e-7 and abs(d[2] - s[2]) <= 2e-6 * abs(s[2]) + 1e-7, (d, s)
        assert d[4] == s[4] and d[6] == s[6]                       # labelled-row counts: exact
        assert abs(d[5] - s[5]) <= 1 and abs(d[7] - s[7]) <= 1     # wrong counts: a borderline row at most
    assert abs(r["dist_test"][0] - r["single_test"][0]) <= 2e-6 * abs(r["single_test"][0]) + 1e-7
    # Adam divides by sqrt(v): a weight whose gradient is ~0 amplifies rounding-level differences of the reduction order
    assert r["w1_maxdiff"] <= 5e-5 * r["w1_scale"] and r["w2_maxdiff"] <= 5e-5 * r["w2_scale"]


@pytest.mark.parametrize("preset,scale,hidden", [("reddit", 0.02, 16), ("products", 0.004, 256)])
def test_partitioned_reupload_equals_single(preset, scale, hidden):
    """set_input_host / epoch_prefetch on a row partition: every rank uploads its own rows (scaled differently every step) and gets
    the single-GPU engine's numbers for the same inputs.  The wide plan (hidden 256) reads every node's features, so its ranks
    all-gather the uploaded slices over NVLink."""
    world = 2
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29650", str(ROOT / "tools" / "dist_check.py"), "--preset", preset, "--scale", str(scale),
           "--epochs", "6", "--dropout", "0.5", "--hidden", str(hidden), "--reupload"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert r["replicated"], "ranks hold different weights"
    assert len({row[0] for row in r["single"]}) == len(r["single"])
    for d, s in zip(r["dist"], r["single"]):
        assert abs(d[0] - s[0]) <= 2e-6 * abs(s[0]) + 1e-7 and abs(d[2] - s[2]) <= 2e-6 * abs(s[2]) + 1e-7, (d, s)
    assert abs(r["dist_test"][0] - r["single_test"][0]) <= 2e-6 * abs(r["single_test"][0]) + 1e-7


def test_cli_multi_gpu_matches_single(tmp_path):
    """`GCN_GPUS=2 ./gcn-cuda <dataset>` (one forked worker per GPU, NCCL id over pipes) prints the same epochs as the
    single-GPU CLI on the same files and seed."""
    import os
    import re
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, str(ROOT))
    from tests.util import make_dataset, write_text_dataset
    gd = make_dataset(n=3000, f=120, c=6, n_undirected=20000, nnz_per_row=9, seed=12, alpha=1.4)
    write_text_dataset(tmp_path, "toy", gd)
    outs = []
    for gpus in ("1", "2"):
        env = dict(os.environ, GCN_SEED="9", GCN_GPUS=gpus, GCN_PLAN="fused")
        r = subprocess.run([str(ROOT / "gcn-cuda"), "toy", "-", "-", "16", "-", "0.5", "0.01", "5e-4", "6"], cwd=tmp_path, env=env,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("epoch=") or l.startswith("test_loss")])
    assert len(outs[0]) == 7 and len(outs[1]) == 7
    for a, b in zip(*outs):
        fa = [float(x) for x in re.findall(r"=(\d+\.\d+)", a)]
        fb = [float(x) for x in re.findall(r"=(\d+\.\d+)", b)]
        assert abs(fa[0] - fb[0]) <= 2e-5 and abs(fa[-3] - fb[-3]) <= 2e-5 if a.startswith("epoch") else abs(fa[0] - fb[0]) <= 2e-5, (a, b)
