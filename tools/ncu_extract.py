#!/usr/bin/env python
"""tools/ncu_extract.py — turn an `ncu --set full` capture (.ncu-rep, read here without a GPU) into the small CSV kept under
profiles/, and (for the GraphSum gather) refresh profiles/graphsum_traffic.json, which bench.py reports as
`roofline.traffic` only while its `graph_cu_sha16` matches the csrc/graph.cu that is built.

    python tools/ncu_extract.py gpurun_out/r02j_prof_gather.ncu-rep profiles/r02j_gather_ncu_full.csv [--traffic gather_kernel]
"""
import csv
import hashlib
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")


def main():
    rep, out = sys.argv[1], Path(sys.argv[2])
    traffic_kernel = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["ID", "Kernel Name"] + [k for k in KEEP if k in idx]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(cols)
        w.writerow(["", ""] + [units[idx[k]] for k in cols[2:]])
        for r in data:
            w.writerow([r[idx["ID"]], r[idx["Kernel Name"]][:120]] + [r[idx[k]] for k in cols[2:]])
    print(f"{out}: {len(data)} launches, {len(cols) - 2} metrics")
    if traffic_kernel:
        sel = [r for r in data if traffic_kernel in r[idx["Kernel Name"]]]
        def num(r, k): return float(r[idx[k]].replace(",", ""))
        def in_bytes(r, k):
            v, u = num(r, k), units[idx[k]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        # the full-graph launches are the long ones: take those within 20 % of the longest
        dur = [num(r, "gpu__time_duration.sum") for r in sel]
        full = [r for r, d in zip(sel, dur) if d >= 0.8 * max(dur)]
        tr = sum(in_bytes(r, "dram__bytes_read.sum") + in_bytes(r, "dram__bytes_write.sum") for r in full) / len(full)
        rec = {"kernel": full[0][idx["Kernel Name"]][:100], "dram_bytes_per_launch": tr, "unit": "bytes", "launches_averaged": len(full),
               "source": f"{out} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean over the full-graph launches of the capture)",
               "graph_cu_sha16": hashlib.sha256((ROOT / "cuda_gcn_b200" / "csrc" / "graph.cu").read_bytes()).hexdigest()[:16]}
        (ROOT / "profiles" / "graphsum_traffic.json").write_text(json.dumps(rec, indent=1))
        print("profiles/graphsum_traffic.json:", rec)


if __name__ == "__main__":
    main()
