#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` report (needs -lineinfo):
    python scripts/ncu_source_lines.py <report.ncu-rep> <kernel regex> [launch index] [top N]
Runs `ncu --page source --csv --print-source cuda,sass` and adds up, per CUDA source line, the warp stall samples
and executed warp instructions of the SASS attributed to it (files of inlined code included)."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", f"regex:{kern}", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur_file, hdr, cur_line, cur_src = None, None, None, None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_samp, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
            i_stalls = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur_line, cur_src = r[0], r[1]
        if len(r) <= i_inst or r[2] in ("-", "..."):
            continue
        key = (cur_file, cur_line, cur_src.strip()[:90])
        a = agg.setdefault(key, {"samples": 0, "inst": 0, "stalls": {}})
        try:
            a["samples"] += int(r[i_samp])
            a["inst"] += int(r[i_inst])
        except ValueError:
            continue
        for h, i in i_stalls.items():
            if i < len(r) and r[i] not in ("", "-", "0"):
                a["stalls"][h] = a["stalls"].get(h, 0) + int(r[i])
    tot_s = sum(a["samples"] for a in agg.values()) or 1
    tot_i = sum(a["inst"] for a in agg.values()) or 1
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    print("| file:line | samples | % | warp inst | top stalls | source |")
    print("|---|---:|---:|---:|---|---|")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{h[6:]} {v}" for h, v in sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print(f"| {key[0]}:{key[1]} | {a['samples']} | {100 * a['samples'] / tot_s:.1f} | {a['inst']} | {st} | `{key[2]}` |")


if __name__ == "__main__":
    main()
