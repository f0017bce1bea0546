#!/usr/bin/env python
"""tools/reddit_preprocess.py — GraphSAGE-format dataset -> the three text files the parser reads.

Does what the reference's one-off notebook export `reddit_preprocess.py` does for the Reddit dataset
(reference reddit_preprocess.py:27-167), re-implemented on numpy only (the reference needs networkx and sklearn):

    <prefix>-G.json          node-link graph: nodes [{"id", "val", "test", ...}], links [{"source", "target"}]
    <prefix>-feats.npy       float features, row id_map[node id]
    <prefix>-id_map.json     node id -> feature row
    <prefix>-class_map.json  node id -> integer class

  * nodes without both a `val` and a `test` annotation are dropped (:52-56);
  * nodes are renumbered by their SORTED id (:102-105); node i's line in <out>.graph lists its neighbours in link order,
    both directions of every undirected link, duplicates collapsed (a simple graph), no self entry (the parser adds it);
  * split codes: 1 = neither val nor test (train), 2 = val, 3 = test (:136-155);
  * features are standardised with the mean / population standard deviation of the TRAINING rows (:71-77; a zero
    deviation is left at 1 as sklearn's StandardScaler does) and written as `label k:v ...` with 0-based keys (:161-167).
    Exactly-zero values are omitted, as sklearn.datasets.dump_svmlight_file does — note that the engine's dense fast
    path needs every row to store all columns, so a dataset with exact zeros after scaling runs the CSR kernels.

`links` may refer to nodes by position in the node list (GraphSAGE's files, written by networkx 1.x) or by id.

    python tools/reddit_preprocess.py --prefix /data/reddit/reddit --out data/reddit
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np


def load(prefix: str):
    g = json.load(open(prefix + "-G.json"))
    nodes, links = g["nodes"], g["links"]
    id_map = {str(k): int(v) for k, v in json.load(open(prefix + "-id_map.json")).items()}
    class_map = json.load(open(prefix + "-class_map.json"))
    feats = np.load(prefix + "-feats.npy")
    return nodes, links, id_map, {str(k): v for k, v in class_map.items()}, feats


def convert(nodes, links, id_map, class_map, feats):
    keep = [i for i, nd in enumerate(nodes) if "val" in nd and "test" in nd]
    dropped = len(nodes) - len(keep)
    names = [nodes[i]["id"] for i in keep]
    # sorted ids define the numbering; python's sort on the ids themselves (ints or strings), as the reference's zip/sorted
    order = sorted(range(len(keep)), key=lambda j: names[j])
    new_id = {names[j]: rank for rank, j in enumerate(order)}
    pos_to_new = {keep[j]: new_id[names[j]] for j in range(len(keep))}     # position in the node list -> new id
    n = len(keep)

    by_position = all(isinstance(l["source"], int) and isinstance(l["target"], int) for l in links[:1000]) and \
        not all(isinstance(nd["id"], int) for nd in nodes[:1000])
    adj = [dict() for _ in range(n)]                                        # insertion-ordered neighbour sets
    for l in links:
        if by_position:
            a, b = pos_to_new.get(l["source"]), pos_to_new.get(l["target"])
        else:
            a, b = new_id.get(l["source"]), new_id.get(l["target"])
        if a is None or b is None:
            continue
        adj[a][b] = True
        adj[b][a] = True

    val = np.array([bool(nodes[keep[j]]["val"]) for j in order])
    test = np.array([bool(nodes[keep[j]]["test"]) for j in order])
    split = np.where(val, 2, np.where(test, 3, 1)).astype(np.int32)          # train, else val, else test (:141-146)
    labels = np.array([int(class_map[str(names[j])]) for j in order], np.int32)
    rows = np.array([id_map[str(names[j])] for j in order], np.int64)

    x = np.asarray(feats, np.float64)
    train_rows = rows[split == 1]
    mean = x[train_rows].mean(axis=0)
    std = x[train_rows].std(axis=0)
    std[std == 0] = 1.0
    x = ((x - mean) / std)[rows]                                            # row i = node i in the new numbering
    return dict(n=n, dropped=dropped, adj=adj, split=split, labels=labels, x=x)


def write_text(out: Path, d):
    out.parent.mkdir(parents=True, exist_ok=True)
    with open(str(out) + ".graph", "w") as fh:
        for nb in d["adj"]:
            fh.write(" ".join(str(v) for v in nb) + "\n")
    with open(str(out) + ".split", "w") as fh:
        fh.write("".join(f"{int(s)}\n" for s in d["split"]))
    with open(str(out) + ".svmlight", "w") as fh:
        x = d["x"]
        for i in range(d["n"]):
            nz = np.nonzero(x[i])[0]
            fh.write(str(int(d["labels"][i])) + "".join(f" {int(k)}:{x[i, k]:.16g}" for k in nz) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prefix", required=True, help="path prefix of the GraphSAGE files (e.g. /data/reddit/reddit)")
    ap.add_argument("--out", required=True, help="output path prefix (e.g. data/reddit)")
    a = ap.parse_args()
    d = convert(*load(a.prefix))
    print(f"{d['n']} nodes ({d['dropped']} dropped), {sum(len(nb) for nb in d['adj'])} directed edges, "
          f"{(d['split'] == 1).sum()} train / {(d['split'] == 2).sum()} val / {(d['split'] == 3).sum()} test", file=sys.stderr)
    write_text(Path(a.out), d)   # the engine's parser writes its .gcnbin cache beside them on the first run


if __name__ == "__main__":
    main()
