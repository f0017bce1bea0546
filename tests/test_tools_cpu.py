"""CPU tests of tools/reddit_preprocess.py (the GraphSAGE -> .graph/.split/.svmlight converter, reference
reddit_preprocess.py:27-167) on a tiny synthetic GraphSAGE-format dataset; the outputs are read back through the
engine's own parser.  The reference script itself does not run under the networkx in this image (it uses the 1.x
`G.node` API), so the expectations are built here from its stated steps, with sklearn's StandardScaler and
dump_svmlight_file (the two library calls it makes) as independent checks where sklearn is importable."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def make_graphsage(tmp, by_position=True, seed=3):
    rng = np.random.default_rng(seed)
    n, f, c = 60, 12, 5
    names = [f"t3_{v:04x}" for v in rng.permutation(4096)[:n]]           # ids whose sorted order differs from file order
    kind = rng.choice(3, n, p=[0.6, 0.2, 0.2])                            # 0 train, 1 val, 2 test
    nodes = [{"id": nm, "val": bool(k == 1), "test": bool(k == 2)} for nm, k in zip(names, kind)]
    nodes.insert(7, {"id": "broken"})                                    # lacks annotations: dropped
    pos = [i for i in range(n + 1) if i != 7]
    links, seen = [], set()
    while len(links) < 150:
        a, b = (int(v) for v in rng.integers(0, n, 2))
        if a == b or (a, b) in seen or (b, a) in seen:
            continue
        seen.add((a, b))
        links.append({"source": pos[a], "target": pos[b]} if by_position else {"source": names[a], "target": names[b]})
    links.append(dict(links[0]))                                          # a duplicate link collapses
    links.append({"source": 7, "target": pos[0]} if by_position else {"source": "broken", "target": names[0]})
    feats = rng.normal(2.0, 3.0, (n + 5, f))
    feats[:, 4] = 1.5                                                     # zero deviation: scale stays 1, values become 0
    rows = rng.permutation(n + 5)[:n]
    id_map = {nm: int(r) for nm, r in zip(names, rows)}
    id_map["broken"] = int(n + 4)
    class_map = {nm: int(v) for nm, v in zip(names, rng.integers(0, c, n))}
    class_map[names[0]] = c - 1
    prefix = str(tmp / "mini")
    json.dump({"directed": False, "graph": {}, "nodes": nodes, "links": links, "multigraph": False}, open(prefix + "-G.json", "w"))
    json.dump(id_map, open(prefix + "-id_map.json", "w"))
    json.dump(class_map, open(prefix + "-class_map.json", "w"))
    np.save(prefix + "-feats.npy", feats)
    return dict(prefix=prefix, names=names, kind=kind, seen=seen, feats=feats, id_map=id_map, class_map=class_map, n=n, f=f, c=c)


@pytest.fixture(scope="module")
def host():
    subprocess.run(["make", "-C", str(ROOT / "cuda_gcn_b200" / "host")], check=True, capture_output=True)
    import os
    os.environ["GCN_NO_CACHE"] = "1"
    from cuda_gcn_b200 import host_api
    host_api.load()
    return host_api


@pytest.mark.parametrize("by_position", [True, False])
def test_reddit_preprocess(host, tmp_path, by_position):
    g = make_graphsage(tmp_path, by_position)
    out = tmp_path / "data" / "mini"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "reddit_preprocess.py"), "--prefix", g["prefix"], "--out", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "60 nodes (1 dropped)" in r.stderr

    d = host.Data.parse(tmp_path / "data", "mini")
    assert d is not None
    a = d.arrays()
    n, f, c, names = g["n"], g["f"], g["c"], g["names"]
    assert (d.params.num_nodes, d.params.output_dim) == (n, c)
    order = sorted(range(n), key=lambda j: names[j])                     # new id -> original index
    rank = {j: i for i, j in enumerate(order)}

    # graph: row i = self loop, then the neighbours of the i-th id in sorted order; symmetric; duplicates collapsed
    want = [set() for _ in range(n)]
    for x, y in g["seen"]:
        want[rank[x]].add(rank[y])
        want[rank[y]].add(rank[x])
    ip, ix = a["graph_indptr"], a["graph_indices"]
    for i in range(n):
        row = ix[ip[i]:ip[i + 1]]
        assert row[0] == i and len(row) == len(want[i]) + 1 and set(row[1:].tolist()) == want[i], i

    # split codes and labels in the new numbering
    assert (a["split"] == np.array([[1, 2, 3][g["kind"][j]] for j in order])).all()
    assert (a["label"] == np.array([g["class_map"][names[j]] for j in order])).all()

    # features: standardised with the training rows' statistics, in sorted-id order; exact zeros are not stored
    rows = np.array([g["id_map"][names[j]] for j in order])
    train = rows[a["split"] == 1]
    mean, std = g["feats"][train].mean(0), g["feats"][train].std(0)
    std[std == 0] = 1
    x = ((g["feats"] - mean) / std)[rows]
    assert np.abs(x[a["split"] == 1].mean(0)).max() < 1e-12
    dense = np.zeros((n, f), np.float32)
    fp, fi, fv = a["feature_indptr"], a["feature_indices"], a["feature_value"]
    for i in range(n):
        dense[i, fi[fp[i]:fp[i + 1]]] = fv[fp[i]:fp[i + 1]]
        assert 4 not in fi[fp[i]:fp[i + 1]]                              # the constant column scaled to exactly 0
    assert (fp[1:] - fp[:-1] == f - 1).all()
    assert (dense.view(np.uint32) == x.astype(np.float32).view(np.uint32)).all() or np.abs(dense - x).max() < 1e-6

    try:
        from sklearn.datasets import dump_svmlight_file
        from sklearn.preprocessing import StandardScaler
    except ImportError:
        return
    sc = StandardScaler().fit(g["feats"][train])
    xs = sc.transform(g["feats"])[rows]
    assert np.abs(xs - x).max() < 1e-12
    ref = tmp_path / "sk.svmlight"
    dump_svmlight_file(xs, a["label"], str(ref))
    # same tokens as sklearn's writer (labels, 0-based keys, skipped zeros); values equal to within the last printed digit
    for mine, theirs in zip(open(str(out) + ".svmlight"), open(ref)):
        tm, tt = mine.split(), theirs.split()
        assert tm[0] == tt[0] and [t.split(":")[0] for t in tm[1:]] == [t.split(":")[0] for t in tt[1:]]
        vm = np.array([float(t.split(":")[1]) for t in tm[1:]]); vt = np.array([float(t.split(":")[1]) for t in tt[1:]])
        assert np.allclose(vm, vt, rtol=1e-13, atol=0)


def test_reddit_preprocess_matches_reference_script(tmp_path):
    """tools/reddit_preprocess.py against the output of the UNMODIFIED reference script on the same GraphSAGE-format input
    (tests/golden/reddit_preprocess/, made by tools/make_reddit_preprocess_golden.py, which executes
    /root/reference/reddit_preprocess.py under three import shims): the .graph and .split files are identical text,
    the .svmlight file has the same labels, the same keys and values equal to 1e-12 relative (StandardScaler vs numpy)."""
    gold = ROOT / "tests" / "golden" / "reddit_preprocess"
    out = tmp_path / "reddit"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "reddit_preprocess.py"), "--prefix", str(gold / "reddit"), "--out", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "reddit.graph").read_text() == (gold / "reddit.graph").read_text()
    assert (tmp_path / "reddit.split").read_text() == (gold / "reddit.split").read_text()
    ours = (tmp_path / "reddit.svmlight").read_text().splitlines()
    ref = (gold / "reddit.svmlight").read_text().splitlines()
    assert len(ours) == len(ref) == 56
    for a, b in zip(ours, ref):
        fa, fb = a.split(), b.split()
        assert int(float(fa[0])) == int(float(fb[0]))                       # sklearn writes the label as given
        ka, kb = [t.split(":") for t in fa[1:]], [t.split(":") for t in fb[1:]]
        assert [k for k, _ in ka] == [k for k, _ in kb]
        for (_, va), (_, vb) in zip(ka, kb):
            assert abs(float(va) - float(vb)) <= 1e-12 * max(1.0, abs(float(vb)))
