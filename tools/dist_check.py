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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="pubmed")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--hidden", type=int, default=16)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    abi.require_device(local)
    data = host_api.Data.synth(a.preset, a.scale)
    uid = host_api.rendezvous(rank, world, os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")))
    eng = host_api.Engine(data, hidden_dim=a.hidden, dropout=a.dropout, seed=a.seed, plan=host_api.PLAN_FUSED, device=local, rank=rank, world=world, nccl_id=uid)
    dist_series, dist_test = series(eng, a.epochs)
    w1, w2 = eng.var(2), eng.var(5)
    # every rank must hold bit-identical replicated weights
    chk = eng.allreduce_host([float(np.abs(w1).sum()), -float(np.abs(w1).sum())], op_max=True)
    replicated = bool(chk[0] == -chk[1])
    eng.close()
    if rank == 0:
        single = host_api.Engine(data, hidden_dim=a.hidden, dropout=a.dropout, seed=a.seed, plan=host_api.PLAN_FUSED, device=local)
        one_series, one_test = series(single, a.epochs)
        v1, v2 = single.var(2), single.var(5)
        single.close()
        print(json.dumps({"world": world, "dist": dist_series, "single": one_series, "dist_test": dist_test, "single_test": one_test,
                          "w1_maxdiff": float(np.abs(w1 - v1).max()), "w1_scale": float(np.abs(v1).max()),
                          "w2_maxdiff": float(np.abs(w2 - v2).max()), "w2_scale": float(np.abs(v2).max()), "replicated": replicated}))


if __name__ == "__main__":
    main()
