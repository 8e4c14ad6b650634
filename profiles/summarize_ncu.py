"""Turns ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/<name>_launches.txt
  python profiles/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep  > profiles/<name>_kernel.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                u = d["Metric Unit"]
                ms = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
                k = re.sub(r"\(.*", "", d["Kernel Name"])[:48]
                agg[k][0] += 1
                agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel device time (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:50s} launches={v[0]:5d}  ms={v[1]:9.3f}  share={v[1] / tot * 100:5.1f}%")
    print(f"{'total':50s} launches={sum(v[0] for v in agg.values()):5d}  ms={tot:9.3f}")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print(f"## {d.get('Kernel Name', '?')}  (ncu --set full --clock-control none)")
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"{h:75s} {vals[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 3:
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr)]

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        tot = sum(f(r, "# Samples") for r in data) or 1.0
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        print("## warp stall reasons (share of samples)")
        for s_, v in sorted(((s_, sum(f(r, s_) for r in data)) for s_ in stalls), key=lambda x: -x[1])[:8]:
            print(f"{s_:30s} {v / tot * 100:5.1f}%")
        op = collections.Counter()
        ti = sum(f(r, "Instructions Executed") for r in data) or 1.0
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            if m:
                op[m.group(2).split(".")[0]] += f(r, "Instructions Executed")
        print("## SASS opcode mix (share of executed warp instructions)")
        print("  ".join(f"{k}={v / ti * 100:.1f}%" for k, v in op.most_common(16)))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
