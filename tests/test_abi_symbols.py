"""CPU-only: the C-ABI libraries load and export every symbol their headers declare (no compute calls)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return sorted(set(re.findall(r"\b(gcn[a-z]*_[a-z0-9_]+)\s*\(", text)))


def build(target_dir):
    subprocess.run(["make", "-C", str(ROOT / "cuda_gcn_b200" / target_dir)], check=True, capture_output=True)


def test_libgcnk_exports_header():
    build("csrc")
    from cuda_gcn_b200 import abi
    L = abi.load()
    names = declared("gcnk.h")
    assert len(names) > 50
    for n in names:
        assert hasattr(L, n), f"libgcnk.so does not export {n}"
    # the ctypes table binds every declared function, and nothing that is not declared
    assert sorted(abi.SIGNATURES) == names
    assert L.gcnk_version() >= 100


def test_no_device_is_loud():
    """Without a GPU the library reports GCNK_ENODEVICE and the Python face raises; nothing falls back."""
    from cuda_gcn_b200 import abi
    if abi.device_count() > 0:
        pytest.skip("a GPU is present")
    n = C.c_int(-1)
    assert abi.load().gcnk_device_count(C.byref(n)) == -4 and n.value == 0
    with pytest.raises(abi.GcnkError):
        abi.require_device(0)


def test_partition_rows_host():
    """gcnk_partition_rows is host-side integer work: cuts are monotone, cover [0,n], and balance nnz."""
    import numpy as np
    from cuda_gcn_b200 import abi
    from tests.util import make_graph
    indptr, _ = make_graph(n=5000, n_undirected=40000, seed=3, alpha=1.5)
    n = len(indptr) - 1
    for parts in (1, 2, 3, 8):
        cuts = np.zeros(parts + 1, np.int32)
        abi.k.gcnk_partition_rows(indptr.ctypes.data, n, parts, cuts.ctypes.data)
        assert cuts[0] == 0 and cuts[-1] == n and (np.diff(cuts) >= 0).all()
        loads = np.diff(indptr[cuts])
        assert loads.sum() == indptr[-1]
        assert loads.max() <= indptr[-1] / parts + 2 * np.diff(indptr).max()       # cuts are rounded up to even rows
        assert (cuts[1:-1] % 2 == 0).all()
