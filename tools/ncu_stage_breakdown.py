#!/usr/bin/env python
"""Warp-instructions and stall samples of the extraction kernel per pipeline stage, from an .ncu-rep captured with
--set full --import-source on (the source page correlates SASS counters with the CUDA lines).

  python tools/ncu_stage_breakdown.py gpurun_out/prof_r01d.ncu-rep 11010000

Stage = a line range of streamz_b200/csrc/frontend.cu (keep STAGES in step with the file); inlined helpers are reported
under their own file (fft_math.cuh = the butterflies of stages 2-4)."""
import csv
import io
import subprocess
import sys

STAGES = {"1 staging": (150, 161), "prefetch (fetch_tile)": (118, 148), "2 stage A": (162, 186), "prefetch call": (187, 187),
          "3 stage B": (188, 199), "4 split + power": (200, 234), "5a mel": (235, 252), "5b ln": (253, 259), "6 dct": (260, 273),
          "7 delta / z-score / store": (274, 335)}


def main():
    rep, windows = sys.argv[1], float(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:extract_kernel"],
                         capture_output=True, text=True).stdout
    cur, si, ii, agg = None, None, None, {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            si, ii = r.index("# Samples"), r.index("Instructions Executed")
        elif ii is not None and len(r) > ii and r[2] == "-" and r[0].isdigit():
            agg[(cur, int(r[0]))] = (int(r[ii]), int(r[si]))
    tot_i = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1

    def group(f, line):
        if f != "frontend.cu":
            return f
        for name, (a, b) in STAGES.items():
            if a <= line <= b:
                return name
        return "frontend.cu (other)"
    g = {}
    for (f, line), (i, s) in agg.items():
        a = g.setdefault(group(f, line), [0, 0])
        a[0] += i
        a[1] += s
    print(f"warp-instructions per window: {tot_i / windows:.1f}")
    for k, (i, s) in sorted(g.items(), key=lambda x: -x[1][0]):
        print(f"{k:30s} {i / windows:7.1f} instr/window  {100 * i / tot_i:5.1f} %   stall samples {100 * s / tot_s:5.1f} %")


if __name__ == "__main__":
    main()
