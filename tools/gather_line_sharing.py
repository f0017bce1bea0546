#!/usr/bin/env python
"""tools/gather_line_sharing.py — CPU-only: how many L1TEX wavefronts of the width-16 GraphSum gather could merge on
the Reddit-shape benchmark graph.  A warp-wide gather instruction reads 8 source rows of 64 bytes; two of them cost
one wavefront instead of two only when they lie in the same 128-byte line (rows 2k and 2k+1).  Counts that for
  * the kernel's lane mapping (the four lanes of group g take entries 4g..4g+3 of a 32-entry chunk, so one instruction
    reads every fourth entry of the sorted row) and for eight consecutive entries per instruction,
  * the generator's node order and an order sorted by class, then by degree.
DESIGN.md 3.1 quotes the result (0 / 0.1 % and 0 / 1.6 %)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["GCN_NO_CACHE"] = "1"
from cuda_gcn_b200 import host_api  # noqa: E402

host_api.load()
data = host_api.Data.synth("reddit", float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)   # keep alive: arrays() are views
a = data.arrays()
indptr, indices, label = a["graph_indptr"].astype(np.int64), a["graph_indices"], a["label"]
n, nnz = len(indptr) - 1, len(indices)
row = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))


def merged_fraction(idx, consecutive):
    pos = np.arange(nnz, dtype=np.int64) - (indptr[row] & ~3)          # chunks start at the row's begin rounded down to 4
    chunk, within = pos // 32, pos % 32
    instr = within // 8 if consecutive else within % 4
    key = (row * 4096 + chunk) * 4 + instr                              # one gather instruction of one warp
    line = idx.astype(np.int64) // 2
    o = np.lexsort((line, key))
    k, l = key[o], line[o]
    new = np.ones(nnz, bool)
    new[1:] = (k[1:] != k[:-1]) | (l[1:] != l[:-1])
    return 1.0 - new.sum() / nnz


print(f"n {n} nnz {nnz}  edges inside a class: {(label[row] == label[indices]).mean():.3f}")
print(f"generator order : kernel mapping {merged_fraction(indices, False):.4f}   consecutive-8 {merged_fraction(indices, True):.4f}")
rank = np.empty(n, np.int64)
rank[np.lexsort((-np.diff(indptr), label))] = np.arange(n)              # by class, hubs first inside a class
relabelled = rank[indices]
relabelled = relabelled[np.lexsort((relabelled, row))]                  # rows re-sorted by the new ids
print(f"class-sorted    : kernel mapping {merged_fraction(relabelled, False):.4f}   consecutive-8 {merged_fraction(relabelled, True):.4f}")
