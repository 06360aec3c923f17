#!/usr/bin/env python
"""Condense the SOURCE page of an Nsight Compute report (compiled with -lineinfo, captured with --import-source on) into
per-source-line totals: executed warp instructions, local-memory loads/stores, sampled stalls.  Run on the GPU box
(the .ncu-rep files are too large to bring back); the text goes under profiles/.

usage: tools/ncu_source_hot.py prof.ncu-rep [kernel-regex] [--top N] > profiles/rNN_<kernel>_source.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 45
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                             capture_output=True, text=True).stdout
    # the page is a sequence of tables (one per kernel / view); keep rows that carry a source location
    # the page is a sequence of sections:  "File Path",<path> / "Function Name",<kernel> / header row / rows; a row with
    # a line number is the total of that source line, the rows below it (empty line number) are its SASS instructions
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, path = None, "?"
    per = collections.defaultdict(lambda: collections.Counter())
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit():
            continue
        c = per[f"{path}:{r[0]}"]
        for col, v in zip(hdr, r):
            if col == "Instructions Executed":
                c["inst"] += num(v)
            elif col == "Warp Stall Sampling (All Samples)":
                c["stall"] += num(v)
            elif col == "L2 Theoretical Sectors Local":
                c["ldl"] += num(v)
            elif col.startswith("stall_") and "Not Issued" not in col:
                c[col] += num(v)
        c["text"] = r[1].strip()[:80]
    tot = collections.Counter()
    for c in per.values():
        for k in ("inst", "stall", "ldl", "stl"):
            tot[k] += c[k]
    print(f"# {rep}: per-source-line totals; executed warp instructions {tot['inst']:.0f}, stall samples {tot['stall']:.0f}, "
          f"local loads {tot['ldl']:.0f}, local stores {tot['stl']:.0f}")
    reasons = collections.Counter()
    for c in per.values():
        for k, v in c.items():
            if k.startswith("stall_"):
                reasons[k] += v
    print("# stall reasons (samples):", ", ".join(f"{k[6:]} {v:.0f}" for k, v in reasons.most_common(10)))
    print(f"{'location':28s} {'inst%':>6s} {'stall%':>6s} {'L2 local sectors':>16s}  top reasons / source")
    for key, c in sorted(per.items(), key=lambda kv: -kv[1]["stall"])[:top]:
        why = ",".join(f"{k[6:]}={v:.0f}" for k, v in sorted(((k, v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3])
        print(f"{key:28s} {100 * c['inst'] / max(tot['inst'], 1):6.2f} {100 * c['stall'] / max(tot['stall'], 1):6.2f} "
              f"{c['ldl']:16.0f}  [{why}] {c['text']}")


if __name__ == "__main__":
    main()
