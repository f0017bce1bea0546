"""CPU-only tests of the C++ host layer (libgcnhost.so): parser against the reference's golden outputs,
the synthetic generator's invariants, exported symbols.  No GPU calls."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def host():
    subprocess.run(["make", "-C", str(ROOT / "cuda_gcn_b200" / "host")], check=True, capture_output=True)
    import os
    os.environ["GCN_NO_CACHE"] = "1"      # the golden directories stay free of .gcnbin files; one test opts back in
    from cuda_gcn_b200 import host_api
    host_api.load()
    return host_api


def test_libgcnhost_exports_header(host):
    text = (ROOT / "include" / "gcn_host.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(gcnh_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 30
    L = host.load()
    for n in names:
        assert hasattr(L, n), n
    assert sorted(host.SIGNATURES) == names


@pytest.mark.parametrize("name", ["toy", "quirks"])
def test_parser_matches_reference_golden(host, name):
    """Bit-exact integer work: CSR, feature index, labels, split and the derived dims equal what the
    UNMODIFIED reference parser produced for the same files (tools/make_golden.py)."""
    want = np.load(GOLD / f"parser_{name}.npz")
    d = host.Data.parse(GOLD / "parser_toy" / "data", name)
    assert d is not None
    got = d.arrays()
    assert (d.params.num_nodes, d.params.input_dim, d.params.output_dim) == tuple(int(want[k]) for k in ("num_nodes", "input_dim", "output_dim"))
    for k in ("graph_indptr", "graph_indices", "feature_indptr", "feature_indices", "label", "split"):
        assert got[k].shape == want[k].shape and (got[k] == want[k]).all(), k
    assert (got["feature_value"].view(np.uint32) == want["feature_value"].view(np.uint32)).all()


def test_parser_matches_oracle_on_generated_text(host, oracle, tmp_path):
    from tests.util import make_dataset, write_text_dataset
    gd = make_dataset(n=400, f=90, c=6, n_undirected=1500, nnz_per_row=9, seed=5, isolated=7, empty_rows=3)
    root = write_text_dataset(tmp_path, "gen", gd, float_fmt="%.9g")
    d = host.Data.parse(root, "gen")
    got, want = d.arrays(), oracle.parse(tmp_path, "gen")
    for k in ("graph_indptr", "graph_indices", "feature_indptr", "feature_indices", "label", "split"):
        assert (got[k] == want[k]).all(), k
    assert (got["feature_value"].view(np.uint32) == want["feature_value"].view(np.uint32)).all()
    # and the round trip reproduces the generator's arrays
    assert (got["graph_indices"] == gd.graph_indices).all() and (got["feature_value"] == gd.feature_value).all()


def test_binary_cache_round_trip(host, tmp_path, monkeypatch):
    """The .gcnbin cache written after a text parse reloads to exactly the same arrays, is ignored when a text file is
    newer, and can be switched off with GCN_NO_CACHE."""
    import os, time
    from tests.util import make_dataset, write_text_dataset
    gd = make_dataset(n=300, f=70, c=5, n_undirected=900, nnz_per_row=7, seed=8, isolated=4)
    root = write_text_dataset(tmp_path, "c", gd, float_fmt="%.9g")
    monkeypatch.delenv("GCN_NO_CACHE", raising=False)
    first = host.Data.parse(root, "c")
    assert (root / "c.gcnbin").exists()
    second = host.Data.parse(root, "c")                                  # served from the cache
    for k, v in first.arrays().items():
        w = second.arrays()[k]
        assert v.shape == w.shape and (v.view(np.uint32) == w.view(np.uint32)).all(), k
    assert (second.params.num_nodes, second.params.input_dim, second.params.output_dim) == \
           (first.params.num_nodes, first.params.input_dim, first.params.output_dim)
    # a newer text file invalidates the cache: change one split value
    lines = (root / "c.split").read_text().splitlines()
    lines[0] = "3" if lines[0] != "3" else "2"
    (root / "c.split").write_text("\n".join(lines) + "\n")
    os.utime(root / "c.split", (time.time() + 5, time.time() + 5))
    third = host.Data.parse(root, "c")
    assert third.arrays()["split"][0] == int(lines[0])
    monkeypatch.setenv("GCN_NO_CACHE", "1")
    (root / "c.gcnbin").unlink()
    host.Data.parse(root, "c")
    assert not (root / "c.gcnbin").exists()


def test_parser_failures(host, tmp_path):
    assert host.Data.parse(tmp_path, "missing") is None                 # "Cannot read input" path (main.cpp:33-36)
    data = tmp_path / "data"
    data.mkdir()
    (data / "bad.graph").write_text("1\n0\n")
    (data / "bad.split").write_text("1\n2\n")
    (data / "bad.svmlight").write_text("0 3:1.0\n1.0 2:1\n")           # label '1.0': the reference reads garbage here
    assert host.Data.parse(data, "bad") is None
    (data / "bad.svmlight").write_text("0 3:1.0\n1 2:1\n")
    (data / "bad.split").write_text("1\n\n")                            # std::stoi throws on a blank line
    assert host.Data.parse(data, "bad") is None


@pytest.mark.parametrize("preset,scale", [("cora", 1.0), ("citeseer", 1.0), ("pubmed", 0.25), ("reddit", 0.002)])
def test_synth_invariants(host, preset, scale):
    d = host.Data.synth(preset, scale)
    a, s = d.arrays(), d.sizes()
    n = s["num_nodes"]
    ip, ix = a["graph_indptr"], a["graph_indices"]
    assert ip[0] == 0 and ip[-1] == len(ix) and (np.diff(ip) >= 1).all()
    assert (ix[ip[:-1]] == np.arange(n)).all()                          # the parser's implicit self loop first
    assert ix.min() >= 0 and ix.max() < n
    rows = np.repeat(np.arange(n), np.diff(ip))
    nb = np.ones(len(ix), bool); nb[ip[:-1]] = False
    keys = rows[nb].astype(np.int64) * n + ix[nb]
    assert (np.diff(keys) > 0).all()                                    # sorted, unique, no explicit self edge
    assert (rows[nb] != ix[nb]).all()
    rev = ix[nb].astype(np.int64) * n + rows[nb]
    assert np.array_equal(np.sort(rev), keys)                           # symmetric
    assert s["max_degree"] <= 46340
    assert len(a["label"]) == n and len(a["split"]) == n
    assert a["label"].min() == 0 and a["label"].max() == d.params.output_dim - 1
    assert a["feature_indices"].max() == d.params.input_dim - 1
    fp = a["feature_indptr"]
    assert fp[-1] == len(a["feature_value"]) and np.isfinite(a["feature_value"]).all()
    if preset == "reddit":
        assert (np.diff(fp) == d.params.input_dim).all()                # dense rows
    # deterministic
    d2 = host.Data.synth(preset, scale)
    assert np.array_equal(d2.arrays()["graph_indices"], ix) and np.array_equal(d2.arrays()["feature_value"], a["feature_value"])


def test_cli_usage_and_missing_input(host, tmp_path):
    """The CLI's argument errors need no GPU: usage line + exit 1 without a dataset (main.cpp:17-27), `Cannot read
    input: <name>` + exit 1 for a dataset that is not there (main.cpp:33-36)."""
    cli = ROOT / "gcn-cuda"
    r = subprocess.run([str(cli)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1
    assert r.stdout.startswith("gcn-cuda graph_name [num_nodes input_dim hidden_dim output_dim dropout learning_rate, weight_decay epochs early_stopping]")
    r = subprocess.run([str(cli), "nope"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "Cannot read input: nope" in r.stderr
    r = subprocess.run([str(cli), "synth:unknown"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "Cannot read input" in r.stderr
    import os
    r = subprocess.run([str(cli), "nope"], capture_output=True, text=True, cwd=tmp_path, env=dict(os.environ, GCN_GPUS="9"))
    assert r.returncode == 1 and "GCN_GPUS" in r.stderr


def test_array_views_keep_the_dataset_alive(host):
    """Data.arrays() returns zero-copy views; they must stay valid after the last explicit reference to the Data object
    is gone (a temporary `Data.synth(...).arrays()` used to dangle)."""
    import gc
    a = host.Data.synth("cora", 1.0).arrays()
    gc.collect()
    churn = [np.full(200_000, 7, np.int32) for _ in range(50)]     # reuse freed heap blocks, if any were freed
    ip = a["graph_indptr"]
    assert ip[0] == 0 and (np.diff(ip) > 0).all() and ip[-1] == len(a["graph_indices"])
    assert a["graph_indices"].min() >= 0 and a["graph_indices"].max() < len(ip) - 1
    assert a["feature_value"].dtype == np.float32 and np.isfinite(a["feature_value"]).all()
    del churn


def _quirky_text_dataset(rng, root, name):
    """A valid-but-untidy dataset in the accepted input language of SURVEY Appendix B: ragged whitespace, tabs, blank
    graph lines (self-loop-only nodes), blank feature lines (label -1, no features), unsorted and repeated neighbours,
    unsorted and repeated feature keys, floats in every spelling strtof/operator>> accepts, a missing final newline
    (the reference drops that last line), split codes outside 1..3."""
    n = int(rng.integers(1, 40))
    f = int(rng.integers(1, 30))
    c = int(rng.integers(1, 6))
    sep = lambda: rng.choice([" ", "  ", "\t", " \t "])
    glines, flines, slines = [], [], []
    for i in range(n):
        if rng.random() < 0.15:
            glines.append(rng.choice(["", " ", "\t"]))
        else:
            nb = rng.integers(0, n, int(rng.integers(1, 9)))                 # unsorted, may repeat, may include i itself
            glines.append(rng.choice(["", " "]) + sep().join(str(int(v)) for v in nb) + rng.choice(["", " ", "\t"]))
        if rng.random() < 0.1:
            flines.append("")
        else:
            toks = []
            for _ in range(int(rng.integers(0, 7))):
                k = int(rng.integers(0, f))
                v = float(rng.normal()) * 10 ** int(rng.integers(-3, 3))
                toks.append(f"{k}:" + rng.choice([f"{v:.6g}", f"{v:.3e}", f"{v:.9f}", str(int(v)), f"{abs(v) % 1:.4f}".lstrip("0") or "0"]))
            flines.append(str(int(rng.integers(0, c))) + "".join(sep() + t for t in toks) + rng.choice(["", " "]))
        slines.append(str(int(rng.choice([0, 1, 2, 3, 3, 2, 1, 7]))))
    end = lambda: "" if rng.random() < 0.3 else "\n"
    root.mkdir(parents=True, exist_ok=True)
    (root / f"{name}.graph").write_text("\n".join(glines) + end())
    (root / f"{name}.svmlight").write_text("\n".join(flines) + end())
    (root / f"{name}.split").write_text("\n".join(slines) + end())


def test_parser_fuzz_against_reference(host, oracle, tmp_path):
    """Bit-exact integer work on untidy inputs: 60 random datasets parse to exactly the arrays the UNMODIFIED reference
    parser produces (oracle/_ref when it is built, the pinned C restatement otherwise)."""
    from oracle.checker import Ref, ref_available
    checker = Ref() if ref_available() else oracle
    rng = np.random.default_rng(20240611)
    for case in range(60):
        name = f"fz{case}"
        _quirky_text_dataset(rng, tmp_path / "data", name)
        want = checker.parse(tmp_path, name)
        d = host.Data.parse(tmp_path / "data", name)
        assert (d is None) == (want is None), name
        if want is None:
            continue
        got = d.arrays()
        assert (d.params.num_nodes, d.params.input_dim, d.params.output_dim) == (want["num_nodes"], want["input_dim"], want["output_dim"]), name
        for k in ("graph_indptr", "graph_indices", "feature_indptr", "feature_indices", "label", "split"):
            assert got[k].shape == want[k].shape and (got[k] == want[k]).all(), (name, k)
        assert (got["feature_value"].view(np.uint32) == want["feature_value"].view(np.uint32)).all(), name


def test_bitsliced_rng_matches_scalar_stream(tmp_path):
    """cuda_gcn_b200/csrc/rng_bitsliced.cuh (the per-thread part of the bit-sliced keep-bit kernels) compiled for the host: 32
    streams per call against the scalar xorshift128+ stream (rand.cpp:17-28) and keep rule (module.cpp:211-216), for
    128 / 1,024 draws per stream, thresholds incl. the dropout-0.5 special case, and ragged ends."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "rng_bs_host"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(root / "tests" / "rng_bitsliced_host.cpp")], check=True, timeout=300)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout[-2000:]
