"""Small seeded synthetic inputs for tests (numpy only; the big shapes come from the C++ generator)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from oracle.checker import GraphData


def make_graph(n, n_undirected, seed, alpha=1.0, isolated=0, hub=None, symmetric=True):
    """CSR with the implicit self loop first in every row, then sorted unique neighbours
    (parser.cpp:31,40).  `isolated` nodes keep only the self loop; `hub=(node, degree)` forces one
    high-degree row; symmetric=False keeps only one direction (directed input)."""
    rng = np.random.default_rng(seed)
    live = n - isolated
    a = np.minimum((live * rng.random(n_undirected) ** alpha).astype(np.int64), live - 1)
    b = rng.integers(0, live, n_undirected)
    if hub is not None:
        node, deg = hub
        others = rng.choice(np.setdiff1d(np.arange(live), [node]), size=min(deg, live - 1), replace=False)
        a = np.concatenate([a, np.full(len(others), node)])
        b = np.concatenate([b, others])
    perm = rng.permutation(n)                 # isolated nodes end up scattered
    a, b = perm[a], perm[b]
    keep = a != b
    a, b = a[keep], b[keep]
    if symmetric:
        src, dst = np.concatenate([a, b]), np.concatenate([b, a])
    else:
        src, dst = a, b
    key = np.unique(src.astype(np.int64) * n + dst)
    src, dst = key // n, key % n
    deg = np.bincount(src, minlength=n) + 1
    indptr = np.zeros(n + 1, np.int64)
    indptr[1:] = np.cumsum(deg)
    indices = np.empty(indptr[-1], np.int64)
    indices[indptr[:-1]] = np.arange(n)
    pos = indptr[:-1] + 1
    # neighbours are already sorted by (src, dst)
    offs = np.arange(len(src)) - np.concatenate([[0], np.cumsum(deg - 1)])[src]
    indices[pos[src] + offs] = dst
    return indptr.astype(np.int32), indices.astype(np.int32)


def make_features(n, f, nnz_per_row, seed, dense=False, empty_rows=0):
    rng = np.random.default_rng(seed + 1000)
    if dense:
        indptr = np.arange(0, (n + 1) * f, f, dtype=np.int64)
        indices = np.tile(np.arange(f), n)
        values = rng.standard_normal(n * f).astype(np.float32)
        return indptr.astype(np.int32), indices.astype(np.int32), values
    counts = np.clip(rng.poisson(nnz_per_row, n), 1, f)
    if empty_rows:
        counts[rng.choice(n, empty_rows, replace=False)] = 0
    indptr = np.zeros(n + 1, np.int64)
    indptr[1:] = np.cumsum(counts)
    indices = np.empty(indptr[-1], np.int64)
    for i in range(n):
        indices[indptr[i]:indptr[i + 1]] = np.sort(rng.choice(f, counts[i], replace=False))
    values = (rng.random(indptr[-1]) + 0.05).astype(np.float32)
    if n and f:                                  # make sure max key + 1 == f (parser.cpp:90)
        if counts[-1] == 0:
            pass
        else:
            indices[indptr[-1] - 1] = f - 1
    return indptr.astype(np.int32), indices.astype(np.int32), values


def make_dataset(n=200, f=64, c=5, n_undirected=600, nnz_per_row=8, seed=0, dense=False, isolated=0, hub=None,
                 alpha=1.0, splits=(0.3, 0.2, 0.3), empty_rows=0, symmetric=True) -> GraphData:
    gp, gi = make_graph(n, n_undirected, seed, alpha=alpha, isolated=isolated, hub=hub, symmetric=symmetric)
    fp, fi, fv = make_features(n, f, nnz_per_row, seed, dense=dense, empty_rows=empty_rows)
    rng = np.random.default_rng(seed + 2000)
    label = rng.integers(0, c, n).astype(np.int32)
    label[0] = c - 1                            # max label + 1 == c (parser.cpp:91)
    u = rng.random(n)
    split = np.zeros(n, np.int32)
    t1, t2, t3 = np.cumsum(splits)
    split[u < t1] = 1
    split[(u >= t1) & (u < t2)] = 2
    split[(u >= t2) & (u < t3)] = 3
    return GraphData(gp, gi, fp, fi, fv, label, split, input_dim=f, output_dim=c)


def write_text_dataset(directory, name, d: GraphData, float_fmt="%.7g"):
    """Write data/<name>.{graph,split,svmlight} as the parser expects: the graph
    file lists neighbours WITHOUT the implicit self loop; every file ends with a newline."""
    root = Path(directory) / "data"
    root.mkdir(parents=True, exist_ok=True)
    with open(root / f"{name}.graph", "w") as fh:
        for i in range(d.num_nodes):
            row = d.graph_indices[d.graph_indptr[i] + 1:d.graph_indptr[i + 1]]
            fh.write(" ".join(str(int(x)) for x in row) + "\n")
    with open(root / f"{name}.split", "w") as fh:
        for s in d.split:
            fh.write(f"{int(s)}\n")
    with open(root / f"{name}.svmlight", "w") as fh:
        for i in range(d.num_nodes):
            lo, hi = d.feature_indptr[i], d.feature_indptr[i + 1]
            kv = " ".join(f"{int(k)}:{float_fmt % float(v)}" for k, v in zip(d.feature_indices[lo:hi], d.feature_value[lo:hi]))
            fh.write(f"{int(d.label[i])}" + (" " + kv if kv else "") + "\n")
    return root
