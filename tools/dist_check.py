#!/usr/bin/env python
"""tools/dist_check.py — run under torchrun (one rank per GPU): the row-partitioned engine against the
single-GPU engine on the same synthetic dataset and seed.  Rank 0 prints one JSON line with both series.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py --preset pubmed --scale 1.0 --epochs 5 --dropout 0.5
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cuda_gcn_b200 import abi, host_api  # noqa: E402


def series(eng, epochs):
    out = []
    for _ in range(epochs):
        tl, ta = eng.train_epoch()
        ct = eng.last_counts()
        vl, va = eng.eval(2)
        cv = eng.last_counts()
        out.append([tl, ta, vl, va, *ct, *cv])
    test = eng.eval(3)
    return out, [*test, *eng.last_counts()]


def series_reupload(eng, epochs, x_rows):
    """The same loop with this rank's feature rows re-uploaded from pinned host memory every step, scaled differently each time
    (a stale or misplaced slice shows): even steps through set_input_host + epoch, odd steps through epoch_prefetch."""
    import ctypes as C
    L = host_api.load()
    bufs = []
    for k in range(epochs + 1):
        ptr = L.gcnh_alloc_pinned(max(len(x_rows), 1))
        view = np.ctypeslib.as_array((C.c_float * max(len(x_rows), 1)).from_address(ptr))
        view[:len(x_rows)] = x_rows * np.float32(1.0 + 0.05 * k)
        bufs.append(ptr)
    out = []
    eng.set_input_host(bufs[0])
    for k in range(epochs):
        if k % 2:
            out.append(list(eng.epoch_prefetch(2, bufs[k + 1])))      # step k on input k; input k+1 uploaded under it
        else:
            out.append(list(eng.epoch(2)))
            eng.set_input_host(bufs[k + 1])
    test = eng.eval(3)
    res = out, [*test, *eng.last_counts()]
    for b in bufs:
        L.gcnh_free_pinned(b)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="pubmed")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--hidden", type=int, default=16)
    ap.add_argument("--reupload", action="store_true", help="re-upload differently scaled features every step (set_input_host / epoch_prefetch)")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    abi.require_device(local)
    data = host_api.Data.synth(a.preset, a.scale)
    uid = host_api.rendezvous(rank, world, os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")))
    eng = host_api.Engine(data, hidden_dim=a.hidden, dropout=a.dropout, seed=a.seed, plan=host_api.PLAN_FUSED, device=local, rank=rank, world=world, nccl_id=uid)
    if a.reupload:
        x_rows = (data.slice(rank, world)[0] if world > 1 else data).arrays()["feature_value"].copy()
        dist_series, dist_test = series_reupload(eng, a.epochs, x_rows)
    else:
        dist_series, dist_test = series(eng, a.epochs)
    w1, w2 = eng.var(2), eng.var(5)
    # every rank must hold bit-identical replicated weights
    chk = eng.allreduce_host([float(np.abs(w1).sum()), -float(np.abs(w1).sum())], op_max=True)
    replicated = bool(chk[0] == -chk[1])
    eng.close()
    if rank == 0:
        single = host_api.Engine(data, hidden_dim=a.hidden, dropout=a.dropout, seed=a.seed, plan=host_api.PLAN_FUSED, device=local)
        one_series, one_test = series_reupload(single, a.epochs, data.arrays()["feature_value"].copy()) if a.reupload else series(single, a.epochs)
        v1, v2 = single.var(2), single.var(5)
        single.close()
        print(json.dumps({"world": world, "dist": dist_series, "single": one_series, "dist_test": dist_test, "single_test": one_test,
                          "w1_maxdiff": float(np.abs(w1 - v1).max()), "w1_scale": float(np.abs(v1).max()),
                          "w2_maxdiff": float(np.abs(w2 - v2).max()), "w2_scale": float(np.abs(v2).max()), "replicated": replicated}))


if __name__ == "__main__":
    main()
