#!/usr/bin/env python
"""tools/make_reddit_preprocess_golden.py — golden vectors for tools/reddit_preprocess.py from the reference's own script.

Builds a tiny GraphSAGE-format dataset (string node ids, as Reddit's; a few nodes without val/test annotations; one
constant feature column), then EXECUTES THE UNMODIFIED /root/reference/reddit_preprocess.py on it and keeps what it
writes (reddit.graph / .split / .svmlight) beside the inputs under tests/golden/reddit_preprocess/.  The script is a
2019 notebook export for networkx 2.0-2.3 and an old scipy; two shims let it run here without touching its text:
  * networkx >= 2.4 dropped the `Graph.node` alias of `Graph.nodes` that the script uses (reddit_preprocess.py:63-67);
  * `scipy.sparse.linalg.eigen.arpack` (imported, never used: reddit_preprocess.py:15) no longer exists;
  * networkx >= 3.4 reads the edge list of a node-link document from the key "edges"; GraphSAGE's files (and networkx at the
    time) use "links" (reddit_preprocess.py:29).
Run in the container that has /root/reference; the outputs are committed so the tests need neither.
"""
import json
import os
import runpy
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden" / "reddit_preprocess"
REF = Path("/root/reference/reddit_preprocess.py")


def make_inputs(out: Path, seed=11, n=60, f=8, classes=5):
    rng = np.random.default_rng(seed)
    ids = ["t3_%05x" % v for v in rng.choice(1 << 20, n, replace=False)]          # string ids, unsorted
    nodes = []
    for i, name in enumerate(ids):
        nd = {"id": name}
        if i % 17 != 5:                                                            # a few nodes lack the annotations
            u = rng.random()
            nd["val"], nd["test"] = bool(0.6 <= u < 0.75), bool(u >= 0.75)
        nodes.append(nd)
    # The script probes `G.nodes()[0]` (reddit_preprocess.py:30): under networkx >= 2 that is a lookup of the node whose id is
    # the integer 0.  With GraphSAGE's own files such nodes exist — position-indexed links of the networkx-1.x format become
    # integer nodes without annotations, which the script then removes (:49-56, "networkx weirdness") — so the golden input
    # carries one: an unannotated node 0.
    nodes.insert(0, {"id": 0})
    links = []
    seen = set()
    while len(links) < 4 * n:
        a, b = (int(v) for v in rng.integers(0, n, 2))
        if a != b and (a, b) not in seen and (b, a) not in seen:
            seen.add((a, b))
            links.append({"source": ids[a], "target": ids[b]})
    feats = rng.standard_normal((n + 3, f)) * rng.random(f) * 3 + rng.standard_normal(f)
    feats[:, 2] = 1.25                                                             # zero variance: StandardScaler leaves it at 0
    id_map = {name: int(v) for name, v in zip(ids, rng.permutation(n + 3)[:n])}
    class_map = {name: int(rng.integers(0, classes)) for name in ids}
    out.mkdir(parents=True, exist_ok=True)
    json.dump({"directed": False, "multigraph": False, "graph": {}, "nodes": nodes, "links": links}, open(out / "reddit-G.json", "w"))
    np.save(out / "reddit-feats.npy", feats)
    json.dump(id_map, open(out / "reddit-id_map.json", "w"))
    json.dump(class_map, open(out / "reddit-class_map.json", "w"))


def run_reference(workdir: Path):
    import networkx as nx
    if not hasattr(nx.Graph, "node"):
        nx.Graph.node = property(lambda self: self.nodes)
    from networkx.readwrite import json_graph
    orig = json_graph.node_link_graph
    if not getattr(orig, "_links_default", False):
        def node_link_graph(data, *a, **kw):
            kw.setdefault("edges", "links")
            return orig(data, *a, **kw)
        node_link_graph._links_default = True
        json_graph.node_link_graph = node_link_graph
    for name in ("scipy.sparse.linalg.eigen", "scipy.sparse.linalg.eigen.arpack"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.eigsh = None
            sys.modules[name] = m
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        runpy.run_path(str(REF), run_name="__main__")
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    make_inputs(OUT)
    run_reference(OUT)
    print(sorted(p.name for p in OUT.iterdir()))
