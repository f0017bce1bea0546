#!/usr/bin/env python
"""bench.py — full-batch GCN train epochs/s on the Reddit-shape synthetic graph (BASELINE.json metric).

One "step" = one epoch as the reference times it: train_epoch() + eval(2) (gcn.cpp:136-140), dropout on.
    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]
N > 1 is launched by torchrun, one rank per GPU (RANK/LOCAL_RANK/WORLD_SIZE from the environment).

Prints ONE JSON line: value = epochs/s with everything resident in HBM; e2e = the same step driven
through the C face with the feature matrix re-uploaded from pinned host memory every step (what
CUDAGCN::set_input does each pass, cuda_gcn.cu:81-83) and the scalars read back; roofline = the GraphSum
gather kernel (algorithmic bytes / CUDA-event duration) against the measured HBM copy peak; cpu_baseline
= the unmodified reference CPU engine (oracle/_ref) on a bounded sample of the same workload.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libgcnref.so, built from
/root/reference by oracle/Makefile; the C restatement when that library is absent) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

# keep NCCL's "NCCL version ..." banner off stdout (the contract is ONE JSON line); INFO/TRACE set by the caller stay
if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "full-batch GCN train epochs/s (Reddit-shape)"
WORKLOAD = "reddit-shape 2-layer GCN (train_epoch + eval per step), hidden 16, dropout 0.5"
UNIT = "epochs/s"


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def graph_data(d):
    from oracle.checker import GraphData
    a = d.arrays()
    return GraphData(a["graph_indptr"], a["graph_indices"], a["feature_indptr"], a["feature_indices"], a["feature_value"],
                     a["label"], a["split"], input_dim=d.params.input_dim, output_dim=d.params.output_dim)


def cpu_reference_run(scale, steps, warmup, full_nnz, host_api):
    """The reference CPU engine on a reddit-shape sample (scale of the node and edge counts); returns
    (epochs/s extrapolated to the full workload by the edge ratio, description)."""
    from oracle.checker import best_checker
    chk = best_checker()
    d = host_api.Data.synth("reddit", scale)
    s = d.sizes()
    ref = chk.gcn(graph_data(d), dropout=0.5, epochs=steps + warmup, seed=1)
    for _ in range(warmup):
        ref.train_epoch(); ref.eval(2)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.train_epoch(); ref.eval(2)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    ref.close()
    ratio = s["graph_nnz"] / full_nnz
    value = (1.0 / dt) * ratio
    sample = (f"{steps} epoch(s) of train_epoch+eval(2) by the {chk.name} CPU engine on reddit-shape at scale {scale:g} "
              f"({s['num_nodes']} nodes, {s['graph_nnz']} graph nnz, dense 602 features): {dt:.2f} s/epoch, scaled to the full "
              f"workload by the graph-nnz ratio {ratio:.4f}")
    return value, dt, chk.name, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cuda_gcn_b200 import host_api
    # the full workload's size without generating it: the generator is deterministic, sizes recorded by a full run
    full_nnz = FULL_REDDIT_NNZ
    scale = args.ref_scale
    value, dt, kind, sample = cpu_reference_run(scale, args.steps, args.warmup, full_nnz, host_api)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "scale": 1.0, "engine": "reference CPU engine (gcn-seq code), 1 thread, bounded sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference" if kind == "reference" else "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# graph nnz (incl. self loops) of synth preset "reddit" at scale 1, seed 4 (deterministic generator)
FULL_REDDIT_NNZ = 114_862_869


def run_ours(args):
    from cuda_gcn_b200 import abi, host_api
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    abi.require_device(local)
    K = abi.k
    t_gen = time.perf_counter()
    data = host_api.Data.synth("reddit", args.scale)       # every rank generates the same dataset (deterministic)
    sizes = data.sizes()
    t_gen = time.perf_counter() - t_gen
    N, nnzA, nnzX = sizes["num_nodes"], sizes["graph_nnz"], sizes["feature_nnz"]
    H, C, F = 16, data.params.output_dim, data.params.input_dim
    uid = host_api.rendezvous(rank, world, os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")))
    eng = host_api.Engine(data, hidden_dim=H, dropout=0.5, seed=1, plan=host_api.PLAN_FUSED, device=local, rank=rank, world=world,
                          nccl_id=uid)
    L = host_api.load()
    if world > 1:
        mine, r0, r1 = data.slice(rank, world)
        ms_ = mine.sizes()
        n_loc, nnzA_loc, nnzX_loc = ms_["num_nodes"], ms_["graph_nnz"], ms_["feature_nnz"]
        x_local = mine.arrays()["feature_value"]
    else:
        n_loc, nnzA_loc, nnzX_loc = N, nnzA, nnzX
        x_local = data.arrays()["feature_value"]

    def step():
        r = eng.epoch(2)                                   # train_epoch + eval(2), one host sync (what GCN::run does per epoch)
        return r[2], r[3]

    def barrier():
        K.gcnk_device_sync()
        eng.allreduce_host([0.0])                          # NCCL all-reduce + sync: all ranks have reached this point
        K.gcnk_device_sync()

    for _ in range(args.warmup):
        step()
    # ---- timed region: K epochs, everything resident.  Only the dominant kernel (the full-graph GraphSum gather) is
    # bracketed by CUDA events inside the timed region — that is what the roofline is computed from, live.  Timing every
    # op costs an event pair per launch (measured: 0.2 ms of a 1.9 ms step at N = 2), so the full per-op breakdown comes
    # from a short extra loop after the timed region (--timers keeps everything on inside it).
    timers_in_region = args.timers
    L.gcnh_timer_reset()
    if timers_in_region:
        L.gcnh_timer_enable_gpu(1)
    else:
        L.gcnh_timer_enable_mask(1 << L.gcnh_timer_slot(b"gather_full"))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = abi.Event(), abi.Event()
    launches0 = abi.load().gcnk_launch_count()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        last = step()
    ev1.record()
    ev1.sync()
    barrier()
    ms = ev0.elapsed_ms(ev1)
    ms = float(eng.allreduce_host([ms], op_max=True)[0])   # max over ranks
    launches = abi.load().gcnk_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    breakdown_steps = args.steps
    live = host_api.timers()                               # gather_full, timed live inside the region
    if not timers_in_region:
        breakdown_steps = min(args.steps, 5)
        L.gcnh_timer_reset()
        L.gcnh_timer_enable_gpu(1)
        for _ in range(breakdown_steps):
            step()
        barrier()
    timers = host_api.timers()
    L.gcnh_timer_enable_gpu(0)
    value = args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel: the GraphSum gather over the whole (local) graph
    peak, peak_src = measured_peak()
    # full-graph gather launches only (each bracketed by its own event pair); the row/column-subset launches are
    # reported in `breakdown` as gather_part
    g_total, g_launches = live.get("gather_full", (0.0, 0))
    b_min = 4 * nnzA_loc + 4 * (n_loc + 1) + 4 * H * (N + n_loc)        # indices + indptr + source read once + rows written once
    t_launch = g_total / max(g_launches, 1)
    achieved = b_min / t_launch / 1e9 if t_launch > 0 else 0.0
    traffic = None
    tp = ROOT / "profiles" / "graphsum_traffic.json"
    if tp.exists() and world == 1:
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "gather_kernel (GraphSum, dim 16, all rows of this rank)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_min, "avg_launch_us": t_launch * 1e6, "launches_timed": g_launches,
                "share_of_step": g_total / (ms * 1e-3) if ms > 0 else None,
                "l2_to_sm_gather_bytes_per_launch": 64 * nnzA_loc + 4 * nnzA_loc}
    # what actually bounds this kernel: every gathered 64-byte row is one wavefront of the SM's L1TEX
    # LSU data pipe, one per clock per SM — reported beside the HBM figure, not instead of it
    try:
        sms = abi.C.c_int(0)
        K.gcnk_device_info(local, abi.C.byref(sms), None, None, None, None)
        mhz = float((clk or {}).get("sm_mhz") or 0) or 1965.0
        floor_us = nnzA_loc / max(sms.value, 1) / mhz
        roofline["on_chip_bound"] = {"unit": "L1TEX LSU data-pipe wavefronts (1 per gathered 64-byte row per SM per clock)", "sms": sms.value,
                                     "sm_mhz": mhz, "floor_us": floor_us, "frac": floor_us / (t_launch * 1e6) if t_launch > 0 else None}
    except Exception:
        pass
    breakdown = {k: {"ms_per_step": v[0] * 1e3 / breakdown_steps, "calls_per_step": v[1] / breakdown_steps} for k, v in timers.items()
                 if k not in ("train", "test")}

    # ---- e2e: the same step through the C face with the (local rows of the) feature matrix uploaded from pinned
    # host memory each step
    pinned = L.gcnh_alloc_pinned(max(nnzX_loc, 1))
    host_view = np.ctypeslib.as_array((abi.C.c_float * max(nnzX_loc, 1)).from_address(pinned))
    host_view[:nnzX_loc] = x_local
    e2e_steps = max(3, min(args.steps, 10))
    eng.set_input_host(pinned); step()                      # warm the copy path
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if args.e2e_prefetch:
            eng.epoch_prefetch(2, pinned)                   # the step on the current input; H2D of the next step's input under it
        else:
            eng.set_input_host(pinned)                      # H2D of this rank's nnz(X) floats on the engine's stream
            step()                                          # train_epoch + eval(2); each reads its scalars back (D2H)
    barrier()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    e2e_dt = float(eng.allreduce_host([e2e_dt], op_max=True)[0])
    L.gcnh_free_pinned(pinned)
    e2e = {"value": 1.0 / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(nnzX * 4), "d2h_bytes_per_step": world * (2 * 16 + 4),
           "steps": e2e_steps, "api": "gcnh_engine_epoch_prefetch (upload of step k+1 under step k; include/gcn_host.h)" if args.e2e_prefetch else
           "gcnh_engine_set_input_host + gcnh_engine_train_epoch + gcnh_engine_eval (include/gcn_host.h)"}
    eng.close()
    if rank != 0:
        return

    # ---- CPU baseline: the reference engine on a bounded sample (rank 0, N=1)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, dt, kind, sample = cpu_reference_run(args.cpu_scale, 1, 0, nnzA, host_api)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference" if kind == "reference" else "port", "sample": sample}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "engine": "fused plan",
                       "scale": args.scale, "nodes": N, "graph_nnz": nnzA, "feature_nnz": nnzX, "features": F, "classes": C,
                       "max_degree": sizes["max_degree"], "l2": "inputs larger than L2 each step (X 561 MB + CSR indices 459 MB streamed per pass); no flush",
                       "parallelism": "1 GPU" if world == 1 else f"{world}-way row partition (nnz-balanced), NCCL all-gather of the gather source + all-reduce of dW",
                       "generate_s": round(t_gen, 1)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "breakdown": breakdown, "final": {"val_loss": last[0], "val_acc": last[1]}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the Reddit-shape node/edge counts (1.0 = the BASELINE config)")
    ap.add_argument("--cpu-scale", type=float, default=0.125, help="sample of the workload the CPU baseline runs")
    ap.add_argument("--ref-scale", type=float, default=0.03125, help="--impl reference: sample of the workload per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-prefetch", action="store_true", help="e2e loop through gcnh_engine_epoch_prefetch (pipelined upload; not yet validated on a GPU)")
    ap.add_argument("--timers", action="store_true", help="keep the per-op CUDA-event timers on inside the timed region at N > 1")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
