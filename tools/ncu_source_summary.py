#!/usr/bin/env python
"""Per-source-line warp-stall summary of one kernel launch in an .ncu-rep captured with
`ncu --set full --import-source on` from a `-lineinfo` build (no GPU needed: reads the report).

  python tools/ncu_source_summary.py gpurun_out/final_prof_projmlp.ncu-rep [launch] [top] > profiles/...

For every CUDA source line: the warp-state samples attributed to the SASS generated from it
(share of all samples of the launch) and the two dominant stall reasons.  Samples are taken per
warp scheduler at a fixed period, so a line's share = the share of warp-time spent waiting on or
issuing that line (inclusive of code inlined into it; the launch total counts every SASS
instruction once) -- in a warp-specialised kernel most of it is the mbarrier wait of a role that
is AHEAD of the pipeline, which is what tells which role bounds the kernel.
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                          "cuda,sass", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True, check=True).stdout
    cur = fn = hdr = None
    scol = 0
    stall = []
    agg = collections.defaultdict(lambda: [0, collections.Counter(), ""])
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            scol = hdr.index("Warp Stall Sampling (All Samples)")
            stall = [(i, h) for i, h in enumerate(hdr)
                     if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) <= scol or not r[0].strip():
            continue
        try:
            c = int(r[scol])
        except ValueError:
            continue
        a = agg[(cur, int(r[0]))]
        a[0] += c
        a[2] = r[1]
        for i, h in stall:
            try:
                a[1][h] += int(r[i])
            except ValueError:
                pass
    # an inlined instruction is listed under its own line AND under the call site, so the launch
    # totals come from the SASS page, where every instruction appears once
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                           "--launch-skip", str(launch), "--launch-count", "1"],
                          capture_output=True, text=True, check=True).stdout
    tot, reasons, shdr = 0, collections.Counter(), None
    for r in csv.reader(io.StringIO(sass)):
        if r and r[0] == "Address":
            shdr = r
            sc = shdr.index("Warp Stall Sampling (All Samples)")
            scols = [(i, h) for i, h in enumerate(shdr)
                     if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if shdr is None or len(r) <= sc:
            continue
        try:
            tot += int(r[sc])
        except ValueError:
            continue
        for i, h in scols:
            try:
                reasons[h] += int(r[i])
            except ValueError:
                pass
    tot = tot or 1
    print(f"# {rep}, launch {launch}: {fn.split('(')[0] if fn else '?'}")
    print(f"# {tot} warp-state samples; stall reasons over the launch: " +
          ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in reasons.most_common(9)))
    print("samples  share  file:line  [top stall reasons]  source     (inclusive: a call site counts "
          "the samples of the code inlined into it)")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join(f"{k[6:]} {v}" for k, v in a[1].most_common(2))
        print(f"{a[0]:7d} {100 * a[0] / tot:5.1f}%  {f}:{ln}  [{why}]  {a[2].strip()[:100]}")


if __name__ == "__main__":
    main()
