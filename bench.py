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
/root/reference by oracle/Makefile; the C restatement when that library is absent) on the host cores, on the SAME
workload at FULL size: one step is ~70 s on one core (the engine is single-threaded), so it runs as many of the
requested steps as fit --ref-budget-s (default 150 s, at least one) and reports the steps it really timed.

Every line carries `parity`: the per-epoch losses/accuracies of this very run (warm-up + timed steps, from a fresh
engine, seed 1) against tests/golden/reddit_full_trajectory.json — the unmodified reference CPU engine on the same
workload and seed (tools/make_trajectory_fixture.py) — and, for N > 1, against the committed 1-GPU trajectory.
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
# --workload products: BASELINE configs[4] (ogbn-products shape, hidden 256: the wide plan with the tcgen05 GEMMs); the
# default (and the headline metric) is the Reddit shape
WORKLOADS = {
    "reddit": {"preset": "reddit", "hidden": 16, "metric": METRIC, "workload": WORKLOAD, "features": 602, "classes": 41},
    "products": {"preset": "products", "hidden": 256, "metric": "full-batch GCN train epochs/s (ogbn-products-shape)",
                 "workload": "ogbn-products-shape 2-layer GCN (train_epoch + eval per step), hidden 256, dropout 0.5", "features": 100, "classes": 47},
}


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


def cpu_reference_run(scale, steps, budget_s, host_api, wl=None):
    """The reference CPU engine (1 thread: it has no threading) on the reddit-shape workload at `scale` (1.0 = the
    BASELINE configuration).  Runs up to `steps` epochs of train_epoch + eval(2) but stops as soon as another epoch would
    overrun `budget_s` (at least one epoch is timed).  Returns measured numbers only — nothing is extrapolated."""
    from oracle.checker import best_checker
    chk = best_checker()
    wl = wl or WORKLOADS["reddit"]
    d = host_api.Data.synth(wl["preset"], scale)
    s = d.sizes()
    ref = chk.gcn(graph_data(d), hidden_dim=wl["hidden"], dropout=0.5, epochs=max(steps, 1), seed=1)
    times, last = [], None
    t_all = time.perf_counter()
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        tr = ref.train_epoch()
        ev = ref.eval(2)
        times.append(time.perf_counter() - t0)
        last = (*tr, *ev)
        if time.perf_counter() - t_all + max(times) > budget_s:
            break
    ref.close()
    dt = sum(times) / len(times)
    sample = (f"{len(times)} epoch(s) of train_epoch+eval(2) by the {chk.name} CPU engine (gcn-seq code, 1 thread) on {wl['preset']}-shape at scale "
              f"{scale:g} ({s['num_nodes']} nodes, {s['graph_nnz']} graph nnz, dense {wl['features']} features, hidden {wl['hidden']}, dropout 0.5): "
              f"{dt:.2f} s/epoch measured")
    return {"value": 1.0 / dt, "s_per_step": dt, "steps": len(times), "kind": chk.name, "sample": sample, "sizes": s, "last": last}


def workload_config(scale, sizes, extra, wl=None):
    wl = wl or WORKLOADS["reddit"]
    cfg = {"workload": wl["workload"], "scale": scale, "nodes": sizes["num_nodes"], "graph_nnz": sizes["graph_nnz"],
           "feature_nnz": sizes["feature_nnz"], "features": wl["features"], "classes": wl["classes"], "hidden": wl["hidden"]}
    cfg.update(extra)
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cuda_gcn_b200 import host_api
    wl = WORKLOADS[args.workload]
    r = cpu_reference_run(args.scale, args.steps, args.ref_budget_s, host_api, wl)
    kind = "reference" if r["kind"] == "reference" else "port"
    line = {"impl": "reference", "metric": wl["metric"], "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": 0, "requested_steps": args.steps, "requested_warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.scale, r["sizes"], {
                "engine": "reference CPU engine (oracle/_ref/libgcnref.so = unmodified /root/reference/src/seq), 1 thread",
                "steps_note": f"a step is ~70 s on one core: as many of the requested steps as fit {args.ref_budget_s:g} s are timed, no warm-up",
                "data_generator": "cuda_gcn_b200/host/synth.cpp through libgcnhost.so (data only; all arithmetic is the reference library's)"}, wl),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": 1, "kind": kind, "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "final": {"val_loss": r["last"][2], "val_acc": r["last"][3]}}
    print(json.dumps(line))


def trajectory_parity(history, fixture_path, loss_rtol=1e-4):
    """history: [(train_loss, train_acc, val_loss, val_acc)] from a fresh engine; fixture: the committed trajectory."""
    try:
        fx = json.loads(Path(fixture_path).read_text())
    except Exception as e:
        return {"against": str(fixture_path), "available": False, "why": str(e)}
    n = min(len(history), len(fx["epochs"]))
    if n == 0:
        return {"against": str(fixture_path), "available": False, "why": "no epochs to compare"}
    rel, acc = 0.0, 0.0
    for got, want in zip(history[:n], fx["epochs"][:n]):
        rel = max(rel, abs(got[0] - want[0]) / abs(want[0]), abs(got[2] - want[2]) / abs(want[2]))
        acc = max(acc, abs(got[1] - want[1]), abs(got[3] - want[3]))
    return {"against": str(Path(fixture_path).relative_to(ROOT)), "engine": fx.get("engine"), "epochs_compared": n,
            "max_rel_loss_diff": rel, "max_abs_acc_diff": acc, "loss_rtol": loss_rtol, "ok": bool(rel <= loss_rtol and acc <= 0.002)}


def source_sha(path):
    import hashlib
    return hashlib.sha256(Path(path).read_bytes()).hexdigest()[:16]


def run_ours(args):
    from cuda_gcn_b200 import abi, host_api
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    abi.require_device(local)
    K = abi.k
    t_gen = time.perf_counter()
    wl = WORKLOADS[args.workload]
    wide = wl["hidden"] != 16
    data = host_api.Data.synth(wl["preset"], args.scale)   # every rank generates the same dataset (deterministic)
    sizes = data.sizes()
    t_gen = time.perf_counter() - t_gen
    N, nnzA, nnzX = sizes["num_nodes"], sizes["graph_nnz"], sizes["feature_nnz"]
    H, C, F = wl["hidden"], data.params.output_dim, data.params.input_dim
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

    history = []                                           # every epoch of this run, from the fresh engine on

    def step(record=True):
        r = eng.epoch(2)                                   # train_epoch + eval(2), one host sync (what GCN::run does per epoch)
        if record:
            history.append(r)
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
        L.gcnh_timer_enable_mask(1 << L.gcnh_timer_slot(b"graphsum_fw" if wide else b"gather_full"))
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
    n_resident = len(history)                              # epochs on the generator's own features (the e2e loop re-uploads the same values)

    # ---- roofline of the dominant kernel: the GraphSum gather over the whole (local) graph
    peak, peak_src = measured_peak()
    # full-graph gather launches only (each bracketed by its own event pair); the row/column-subset launches are
    # reported in `breakdown` as gather_part
    g_total, g_launches = live.get("graphsum_fw" if wide else "gather_full", (0.0, 0))
    gw = F if wide else H                                                # width of the dominant gather: the input width in the wide plan
    b_min = 4 * nnzA_loc + 4 * (n_loc + 1) + 4 * gw * (N + n_loc)       # indices + indptr + source read once + rows written once
    if wide:
        b_min += 4 * F * 2 * N + N * F // 8                              # + the dropout/pre-scale pass in the same timer: X read, source written, keep bits
    t_launch = g_total / max(g_launches, 1)
    achieved = b_min / t_launch / 1e9 if t_launch > 0 else 0.0
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture — only if that capture was taken from
    # the graph.cu that is built now (the record carries the source hash), otherwise null rather than a stale number
    traffic, traffic_note = None, None
    tp = ROOT / "profiles" / "graphsum_traffic.json"
    if tp.exists() and world == 1:
        try:
            rec = json.loads(tp.read_text())
            if rec.get("graph_cu_sha16") == source_sha(ROOT / "cuda_gcn_b200" / "csrc" / "graph.cu"):
                traffic = rec.get("dram_bytes_per_launch")
                traffic_note = rec.get("source")
            else:
                traffic_note = "profiles/graphsum_traffic.json was captured from a different csrc/graph.cu; not reported"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": ("drop_scale_rows + gather_kernel (GraphSum at the input width 100: A_hat*drop(X), all rows of this rank)" if wide
                                           else "gather_kernel (GraphSum, dim 16, all rows of this rank)"), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_min, "avg_launch_us": t_launch * 1e6, "launches_timed": g_launches,
                "share_of_step": g_total / (ms * 1e-3) if ms > 0 else None,
                "l2_to_sm_gather_bytes_per_launch": 4 * gw * nnzA_loc + 4 * nnzA_loc}
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
    prefetch = not args.e2e_serial
    eng.set_input_host(pinned); step(False)                 # warm the copy path
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if prefetch:
            eng.epoch_prefetch(2, pinned)                   # the step on the current input; H2D of the next step's input under it
        else:
            eng.set_input_host(pinned)                      # H2D of this rank's nnz(X) floats on the engine's stream
            step(False)                                     # train_epoch + eval(2); each reads its scalars back (D2H)
    barrier()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    e2e_dt = float(eng.allreduce_host([e2e_dt], op_max=True)[0])
    L.gcnh_free_pinned(pinned)
    e2e = {"value": 1.0 / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(nnzX * 4), "d2h_bytes_per_step": world * (2 * 16 + 4),
           "steps": e2e_steps, "api": "gcnh_engine_epoch_prefetch (upload of step k+1 under step k; include/gcn_host.h)" if prefetch else
           "gcnh_engine_set_input_host + gcnh_engine_train_epoch + gcnh_engine_eval (include/gcn_host.h)"}
    if wide and world > 1:
        e2e["note"] = ("the row-partitioned wide plan reads every node's features: each rank uploads its own rows and the slices are "
                       "all-gathered over NVLink (NCCL) on the copy stream")
    eng.close()
    if rank != 0:
        return

    # ---- parity of THIS run: every epoch since the engine was created against the reference CPU engine's trajectory
    # on the same workload and seed, and (N > 1) against the committed single-GPU trajectory
    parity = None
    if args.scale == 1.0:
        parity = trajectory_parity(history[:n_resident], ROOT / "tests" / "golden" / f"{wl['preset']}_full_trajectory.json")
        if world > 1:
            one = trajectory_parity(history[:n_resident], ROOT / "tests" / "golden" / f"{wl['preset']}_full_trajectory_gpu1.json", loss_rtol=2e-6)
            parity["parity_vs_single_gpu_rel"] = one.get("max_rel_loss_diff")
            parity["vs_single_gpu"] = one
    if args.save_trajectory:
        Path(args.save_trajectory).write_text(json.dumps({
            "generator": "bench.py --save-trajectory (this engine, fused plan, 1 GPU)" if world == 1 else f"bench.py, {world} GPUs",
            "engine": f"cuda_gcn_b200 fused plan, {world} GPU(s)", "preset": wl["preset"], "scale": args.scale, "seed": 1, "hidden": H,
            "dropout": 0.5, "nodes": N, "graph_nnz": nnzA, "columns": ["train_loss", "train_acc", "val_loss", "val_acc"],
            "epochs": [list(r) for r in history[:n_resident]]}, indent=1))

    # ---- GraphSum per call at each width the reference's models use (SURVEY 8d metric 2): forward == backward launch
    dims = None
    if world == 1 and not args.no_dims and not wide:
        dims = graphsum_dims(abi, data, peak)

    # ---- CPU baseline: the reference engine on the same workload (rank 0, N=1): ONE full-size epoch (~70 s, 1 core)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(args.cpu_scale, 1, 0, host_api, wl)
        cpu = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "reference" if r["kind"] == "reference" else "port",
               "sample": r["sample"], "sample_scale": args.cpu_scale}

    line = {"metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.scale, sizes, {
                "engine": "fused plan (wide: host/gcn_wide.cpp)" if wide else "fused plan", "max_degree": sizes["max_degree"],
                "l2": ("inputs larger than L2 each step (X 980 MB + CSR indices 500 MB streamed per pass); no flush" if wide else
                       "inputs larger than L2 each step (X 561 MB + CSR indices 459 MB streamed per pass); no flush"),
                "parallelism": "1 GPU" if world == 1 else
                f"{world}-way row partition (nnz-balanced); gather sources exchanged by a hand-written push over NVLink peer memory "
                "(CUDA IPC) with per-rank arrival flags checked inside the consuming GraphSum kernel; dW summed by a peer-memory "
                "all-reduce in rank order; NCCL only for setup (handle exchange) and as fallback (GCN_COMM=nccl)",
                "generate_s": round(t_gen, 1)}, wl),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "graphsum_dims": dims, "cpu_baseline": cpu,
            "host_enqueue_ms_per_step": (live["host_enqueue"][0] * 1e3 / max(live["host_enqueue"][1], 1)) if "host_enqueue" in live else None,
            "parity": parity, "breakdown": breakdown, "final": {"val_loss": last[0], "val_acc": last[1]}}
    print(json.dumps(line))


def graphsum_dims(abi, data, peak, dims=(16, 41, 47, 256), reps=5):
    """gcnk_graphsum (= what CUDAGraphSum::forward / backward launch, cuda_module.cu:74-101: the same loop for both
    directions on a symmetric graph) timed per call at each width the reference's models use."""
    K = abi.k
    a = data.arrays()
    indptr, indices = a["graph_indptr"], a["graph_indices"]
    n, nnz = len(indptr) - 1, len(indices)
    g = abi.Graph(indptr, indices)
    out = []
    for dim in dims:
        x = abi.dev(np.random.default_rng(dim).standard_normal((n, dim)).astype(np.float32))
        y = abi.DeviceArray((n, dim), np.float32)
        pitch = (dim + 3) // 4 * 4                                 # the fused plans keep class-width sources at a 16-byte pitch (47 -> 48)
        xs = abi.DeviceArray((n, pitch), np.float32)
        ys = abi.DeviceArray((n, pitch), np.float32)
        K.gcnk_memset(xs.ptr, 0, 4 * n * pitch, None)
        K.gcnk_scale_rows(g.dinv_ptr(), xs.ptr, xs.ptr, n, pitch, None)
        for _ in range(2):
            K.gcnk_graphsum(g.h, x.ptr, y.ptr, dim, None)
            K.gcnk_gather_plain(g.h, xs.ptr, ys.ptr, pitch, None)
        e0, e1, e2 = abi.Event(), abi.Event(), abi.Event()
        K.gcnk_device_sync()
        e0.record()
        for _ in range(reps):
            K.gcnk_graphsum(g.h, x.ptr, y.ptr, dim, None)          # module-level call: pre-scale pass + gather
        e1.record()
        for _ in range(reps):
            K.gcnk_gather_plain(g.h, xs.ptr, ys.ptr, pitch, None)  # the gather alone (what the fused plans launch)
        e2.record(); e2.sync()
        us_call, us_gather = e0.elapsed_ms(e1) / reps * 1e3, e1.elapsed_ms(e2) / reps * 1e3
        b_min = 4 * nnz + 4 * (n + 1) + 8 * n * dim
        out.append({"dim": dim, "gather_only_pitch": pitch, "fw_us": us_call, "bw_us": us_call, "gather_only_us": us_gather, "algorithmic_bytes": b_min,
                    "GBps": b_min / us_call / 1e3, "frac": b_min / us_call / 1e3 / peak,
                    "gather_only_GBps": b_min / us_gather / 1e3, "gather_only_frac": b_min / us_gather / 1e3 / peak})
        del x, y, xs, ys
    K.gcnk_graph_release_scratch(g.h)
    return {"note": "forward and backward are the same launch (module.cpp:103-119 reuses the forward loop on a symmetric graph); "
                    "fw_us/bw_us = gcnk_graphsum (pre-scale pass + gather; widths that are not a multiple of 4 go through rows padded to a "
                    "16-byte pitch), gather_only = gcnk_gather_plain on a pre-scaled source at that pitch",
            "per_dim": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reddit", choices=sorted(WORKLOADS), help="reddit = the headline metric; products = BASELINE configs[4]")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the Reddit-shape node/edge counts (1.0 = the BASELINE config)")
    ap.add_argument("--cpu-scale", type=float, default=1.0, help="workload scale of the cpu_baseline leg (1.0 = the real workload: ~70 s)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: wall budget for the timed full-size epochs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dims", action="store_true", help="skip the per-width GraphSum timing")
    ap.add_argument("--save-trajectory", default=None, help="write this run's per-epoch losses/accuracies (JSON) to this path")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e loop as set_input_host + epoch (upload, then compute) instead of the pipelined "
                    "gcnh_engine_epoch_prefetch (upload of step k+1 on a copy stream under step k; tests/test_gpu_train.py::test_epoch_prefetch_equals_serial)")
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
