#!/usr/bin/env python
"""tools/make_trajectory_fixture.py — per-epoch loss/accuracy trajectory of the UNMODIFIED reference CPU engine
(oracle/_ref/libgcnref.so = /root/reference/src/seq compiled by oracle/Makefile) on a synthetic preset at FULL size.

    python tools/make_trajectory_fixture.py reddit 1.0 30 tests/golden/reddit_full_trajectory.json [hidden] [dropout]

One epoch = train_epoch() + eval(2), exactly what GCN::run does (src/seq/gcn.cpp:136-138); after the last epoch
eval(3) is recorded too (gcn.cpp:154-157).  seed 1 (the seed bench.py and the tests use), dropout from the shared
xorshift128+ stream.  The file is what bench.py's `parity` field and tests/test_gpu_train.py compare the CUDA
engine with at the headline size; it is written after every epoch so a partial run is still usable.
Runs for about 80 s per epoch at Reddit shape on one core (the reference engine is single-threaded).
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    preset, scale, epochs, out = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), Path(sys.argv[4])
    hidden = int(sys.argv[5]) if len(sys.argv) > 5 else 16
    dropout = float(sys.argv[6]) if len(sys.argv) > 6 else 0.5
    from cuda_gcn_b200 import host_api
    from oracle.checker import GraphData, Ref
    d = host_api.Data.synth(preset, scale)
    a = d.arrays()
    s = d.sizes()
    gd = GraphData(a["graph_indptr"], a["graph_indices"], a["feature_indptr"], a["feature_indices"], a["feature_value"],
                   a["label"], a["split"], input_dim=d.params.input_dim, output_dim=d.params.output_dim)
    ref = Ref().gcn(gd, hidden_dim=hidden, dropout=dropout, epochs=epochs, seed=1)
    doc = {"generator": "tools/make_trajectory_fixture.py", "engine": "oracle/_ref/libgcnref.so (unmodified /root/reference/src/seq)",
           "preset": preset, "scale": scale, "seed": 1, "hidden": hidden, "dropout": dropout, "lr": 0.01, "weight_decay": 5e-4,
           "nodes": s["num_nodes"], "graph_nnz": s["graph_nnz"], "feature_nnz": s["feature_nnz"],
           "columns": ["train_loss", "train_acc", "val_loss", "val_acc"], "epochs": [], "seconds_per_epoch": []}
    for e in range(epochs):
        t0 = time.perf_counter()
        row = [*ref.train_epoch(), *ref.eval(2)]
        doc["epochs"].append([float(x) for x in row])
        doc["seconds_per_epoch"].append(round(time.perf_counter() - t0, 2))
        out.write_text(json.dumps(doc, indent=1))
        print(e + 1, row, doc["seconds_per_epoch"][-1], flush=True)
    doc["test_after_last_epoch"] = [float(x) for x in ref.eval(3)]
    out.write_text(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
