#!/usr/bin/env python
"""Turn the scratch ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

usage: tools/make_profiles.py <round-tag> [--launches gpurun_out/launches.csv] [--rep name=gpurun_out/x.ncu-rep ...]
                                            [--summary name=gpurun_out/ncu_x.txt,gpurun_out/ncu_x.json ...]

--summary takes what `tools/ncu_summary.py <rep> --json <json> > <txt>` wrote ON THE GPU BOX (the .ncu-rep files are
too big for gpurun_out's merge-back limit, so they are condensed there).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            agg.setdefault(re.sub(r"\(.*", "", r[ki]), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (per-launch device time; cold-cache and\n"
                 "# serialised: compare SHARES, not absolutes)\n")
        fh.write(f"# source: {os.path.basename(path)}; total {tot/1e3:.1f} us over {sum(len(v) for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write(f"{sum(v)/tot*100:6.2f}%  n={len(v):4d}  avg {sum(v)/len(v)/1e3:9.2f} us  {k}\n")


def _mb(v):
    return float(v[0].replace(",", "")) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[v[1]]


def record_traffic(traffic, name, tag, json_path):
    """profiles/ncu_traffic.json: per captured kernel the DRAM bytes per launch, the FP-pipe fraction (what binds the
    arithmetic-heavy kernels) and the capture's own duration; bench.py reads it for `roofline`."""
    for d in json.load(open(json_path)):
        if "dram__bytes_read.sum" in d:
            key = re.sub(r"void |unnamed>::|\(.*", "", d["kernel"]).strip()
            is64 = "<double" in key
            pipe = d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active" if is64 else
                         "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
            traffic[f"{name}:{key}"] = {
                "dram_mb_per_launch": round(_mb(d["dram__bytes_read.sum"]) + _mb(d["dram__bytes_write.sum"]), 3),
                "fp_pipe_frac": round(float(pipe[0].replace(",", "")) / 100.0, 4) if pipe else None,
                "fp_pipe": "fp64" if is64 else "fma (fp32)",
                "ncu_us": round(float(d["gpu__time_duration.sum"][0].replace(",", "")) *
                                {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[d["gpu__time_duration.sum"][1]], 3), "tag": tag}


def main():
    tag = sys.argv[1]
    args = sys.argv[2:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    i = 0
    while i < len(args):
        if args[i] == "--launches":
            launches(args[i + 1], os.path.join(ROOT, "profiles", f"{tag}_launches.txt"))
            i += 2
        elif args[i] == "--rep":
            name, rep = args[i + 1].split("=")
            old = sys.argv
            buf = io.StringIO()
            sys.argv, so = ["ncu_summary", rep, "--json", "/tmp/_ncu.json"], sys.stdout
            sys.stdout = buf
            try:
                ncu_summary.main()
            finally:
                sys.stdout, sys.argv = so, old
            with open(os.path.join(ROOT, "profiles", f"{tag}_{name}.txt"), "w") as fh:
                fh.write(f"# condensed from {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on)\n")
                fh.write(buf.getvalue())
            record_traffic(traffic, name, tag, "/tmp/_ncu.json")
            i += 2
        elif args[i] == "--summary":
            name, paths = args[i + 1].split("=")
            txt, js = paths.split(",")
            with open(os.path.join(ROOT, "profiles", f"{tag}_{name}.txt"), "w") as fh:
                fh.write("# condensed on the GPU box by tools/ncu_summary.py from one ncu --set full --clock-control none "
                         "--import-source on capture\n")
                fh.write(open(txt).read())
            record_traffic(traffic, name, tag, js)
            i += 2
        else:
            raise SystemExit(f"bad arg {args[i]}")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
