#!/usr/bin/env python
"""tools/dist_cpu_check.py — CPU-only check of the host side of the row-partitioned path, run as world_size
ranks over gloo (no GPU, no NCCL communicator): partition slices re-concatenate to the original dataset bit
for bit, the per-rank offsets into the shared dropout stream are consistent, and the loopback rendezvous hands
every rank the same 128 bytes."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist          # imported before the native libraries on purpose (its NCCL comes first)

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cuda_gcn_b200 import host_api  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    d = host_api.Data.synth("citeseer", 1.0)
    full = {k: v.copy() for k, v in d.arrays().items()}
    sl, r0, r1 = d.slice(rank, world)
    mine = {k: v.copy() for k, v in sl.arrays().items()}
    mine["range"] = (r0, r1)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    uid = host_api.rendezvous(rank, world, os.environ["MASTER_ADDR"], int(os.environ["MASTER_PORT"]), uid=bytes(range(128)))
    uids = [None] * world
    dist.all_gather_object(uids, uid)
    ok = True
    if rank == 0:
        cuts = [g["range"] for g in gathered]
        ok &= cuts[0][0] == 0 and cuts[-1][1] == len(full["label"]) and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        for key in ("graph_indices", "feature_indices", "feature_value", "label", "split"):
            cat = np.concatenate([g[key] for g in gathered])
            ok &= cat.dtype == full[key].dtype and cat.shape == full[key].shape and bool((cat.view(np.uint32) == full[key].view(np.uint32)).all())
        for key in ("graph_indptr", "feature_indptr"):
            off, parts = 0, [np.zeros(1, np.int32)]
            for g in gathered:
                parts.append(g[key][1:] + off)
                off += int(g[key][-1])
            ok &= bool((np.concatenate(parts) == full[key]).all())
        # offsets into the shared dropout stream: rank k's first X draw is feature_indptr[r0] (engine: x_off)
        off = 0
        for g in gathered:
            ok &= int(full["feature_indptr"][g["range"][0]]) == off
            off += len(g["feature_value"])
        # nnz balance of the graph partition
        loads = [len(g["graph_indices"]) for g in gathered]
        ok &= max(loads) <= len(full["graph_indices"]) / world + 2 * int(np.diff(full["graph_indptr"]).max())   # cuts are rounded up to even rows
        ok &= all(u == bytes(range(128)) for u in uids)
        print("DIST_CPU_OK" if ok else "DIST_CPU_FAIL", cuts, loads)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
