#!/usr/bin/env python
"""Text summary of an .ncu-rep for profiles/: the headline 'details' lines, raw counters, the warp-stall breakdown and
the dynamic instruction mix of every captured kernel.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep "header line" > profiles/rNN_xxx_ncu_summary.txt
  python tools/ncu_summary.py --launches gpurun_out/launches.csv "header line" > profiles/rNN_launches.csv
  python tools/ncu_summary.py --traffic gpurun_out/prof.ncu-rep extract_kernel 11010000 "source note"   (prints a JSON entry)
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

DETAILS = ("Duration", "DRAM Throughput", "Memory Throughput", "Mem Busy", "Max Bandwidth", "Compute (SM) Throughput",
           "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler",
           "L1/TEX Hit Rate", "L2 Hit Rate", "Block Size", "Grid Size", "Registers Per Thread", "Dynamic Shared Memory Per Block",
           "Theoretical Occupancy", "Achieved Occupancy")
RAW = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active")


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def summary(rep, header):
    print("# " + header)
    head, units, rows = raw_rows(rep)
    details = ncu("-i", rep, "--page", "details")
    blocks = re.split(r"\n(?=  \S.*\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\))", details)
    for i, v in enumerate(rows):
        name = v[head.index("Kernel Name")]
        print(f"\n== {name} ==")
        blk = next((b for b in blocks if name.split("(")[0].strip().split()[-1] in b), "")
        seen = set()
        for line in blk.splitlines():
            t = line.strip()
            for k in DETAILS:
                if t.startswith(k + " ") and k not in seen:
                    seen.add(k)
                    print("  " + re.sub(r"\s{2,}", "  ", t))
        print("  -- raw counters")
        for k in RAW:
            if k in head:
                print(f"  {k:90s} {v[head.index(k)]} {units[head.index(k)]}")
        for j, k in enumerate(head):                         # tensor pipe (tcgen05) and TMEM / uniform-datapath counters, whatever this ncu names them
            if ("pipe_tensor" in k or "pipe_tmem" in k) and ".avg" in k and "Triage" not in k and k not in RAW and v[j] not in ("", "0", "n/a"):
                print(f"  {k:90s} {v[j]} {units[j]}")
        print("  -- warp stalls per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active)")
        st = [(float(v[j]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for j, h in enumerate(head)
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[j]]
        for val, nm in sorted(st, reverse=True)[:8]:
            print(f"  {nm:28s} {val:6.3f}")
    for v in rows:
        full = v[head.index("Kernel Name")]
        short = full.split("(")[0].replace("void ", "").split("<")[0].strip()
        src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + short))))
        starts = [i for i, r in enumerate(src) if r and r[0] == "Address"]
        if not starts:
            continue
        s0 = starts[0]
        hdr = src[s0]
        end = starts[1] - 1 if len(starts) > 1 else len(src)
        ci, cs, cw, cx = (hdr.index(k) for k in ("Instructions Executed", "Warp Stall Sampling (All Samples)", "L1 Wavefronts Shared", "Source"))
        by, stall, wf = collections.Counter(), collections.Counter(), collections.Counter()
        for r in src[s0 + 1:end]:
            if len(r) <= ci or not r[ci].isdigit():
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[cx].strip())
            op = m.group(2) if m else r[cx].strip()
            by[op] += int(r[ci]); stall[op] += int(r[cs] or 0); wf[op] += int(r[cw] or 0)
        tot, ts = sum(by.values()), max(1, sum(stall.values()))
        print(f"\n== dynamic instruction mix: {short} ==\n  warp instructions {tot}, shared-memory wavefronts {sum(wf.values())}, stall samples {ts}")
        for op, c in by.most_common(20):
            print(f"  {op:10s} {100 * c / tot:5.1f} % of instructions  {100 * stall[op] / ts:5.1f} % of stall samples")


def launches(path, header):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    print("# " + header)
    print("# (cold-cache, serialised launches: compare SHARES, not absolutes)")
    print("id,kernel,duration_ns")
    other = collections.Counter()
    for r in rows:
        name, val = r[4], float(r[-1])
        unit = r[-2]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        if "szb::" in name or "tc::" in name or "_kernel" in name.split("(")[0] and "at::" not in name:
            print(f"{r[0]},{name.split('(')[0].replace('void ', '')},{ns:.0f}")
        else:
            other["torch / other kernels (synthetic data generation, outside every timed region)"] += ns
    for k, v in other.items():
        print(f"-,{k},{v:.0f}")


def traffic(rep, kernel, windows, source):
    head, units, rows = raw_rows(rep)
    for v in rows:
        if kernel in v[head.index("Kernel Name")]:
            def b(k):
                x, u = float(v[head.index(k)]), units[head.index(k)]
                return int(x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u])
            print(json.dumps({kernel: {"dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
                                       "windows_per_launch": int(windows), "source": source}}, indent=1))
            return


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "--traffic":
        traffic(*sys.argv[2:6])
    else:
        summary(sys.argv[1], sys.argv[2])
