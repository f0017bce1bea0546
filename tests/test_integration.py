"""The drop-in boundary, compiled: the reference's UNMODIFIED src/main.cpp (GPU branch, -D__NVCC__) built against
cuda_gcn_b200/host/*.h + libgcnhost.so/libgcnk.so by oracle/Makefile (-> oracle/_ref/gcn-cuda-refmain).
CPU part: it compiles, links and prints the reference's usage line.  GPU part: on a toy dataset it prints the same
lines as ./gcn-cuda (this repo's own main.cpp), which test_gpu_train.py compares with the unmodified gcn-seq."""
import os
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFMAIN = ROOT / "oracle" / "_ref" / "gcn-cuda-refmain"


def test_reference_main_compiles_against_host_headers():
    if Path("/root/reference/src/main.cpp").exists():
        subprocess.run(["make", "-C", str(ROOT / "cuda_gcn_b200" / "host")], check=True, capture_output=True)
        if REFMAIN.exists():
            REFMAIN.unlink()                      # force a rebuild from the sources where they lie
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True, capture_output=True)
    if not REFMAIN.exists():
        pytest.skip("oracle/_ref/gcn-cuda-refmain not built (needs /root/reference)")
    out = subprocess.run([str(REFMAIN)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1                    # main.cpp:17-27: usage, EXIT_FAILURE
    assert out.stdout.startswith("gcn-cuda graph_name [num_nodes input_dim hidden_dim")


@pytest.mark.gpu
def test_reference_main_runs_on_this_engine(tmp_path):
    if not REFMAIN.exists():
        pytest.skip("oracle/_ref/gcn-cuda-refmain did not travel with the snapshot")
    from tests.util import make_dataset, write_text_dataset
    gd = make_dataset(n=300, f=50, c=4, n_undirected=900, nnz_per_row=6, seed=9)
    write_text_dataset(tmp_path, "toy", gd)
    env = dict(os.environ, GCN_SEED="5", GCN_NO_CACHE="1")
    outs = []
    for exe in (REFMAIN, ROOT / "gcn-cuda"):
        r = subprocess.run([str(exe), "toy"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.splitlines())
    a, b = outs
    assert a[:4] == ["Parse Graph Succeeded.", "Parse Node Succeeded.", "Parse Split Succeeded.", "RUNNING ON GPU"]
    strip = lambda l: re.sub(r" ?time=\d+\.\d+", "", l)
    ea = [strip(l) for l in a if l.startswith("epoch=") or l.startswith("test_loss")]
    eb = [strip(l) for l in b if l.startswith("epoch=") or l.startswith("test_loss")]
    assert len(ea) == 101 and ea == eb            # 100 default epochs (gcn.cpp:9-11) + the test line, same engine => same digits
