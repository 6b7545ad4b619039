#!/usr/bin/env python
"""Hot source lines of one kernel from an ncu capture taken with --import-source on:
usage: python scripts/ncu_hot_lines.py REP KERNEL_REGEX [top]
Prints, per CUDA source line, the warp-stall samples, executed warp instructions, L1 tag requests and shared
wavefronts summed over the SASS instructions the line maps to."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, cur_line, cur_src = None, None, None, None
agg = {}
seen_func = 0
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
        col = {name: k for k, name in enumerate(hdr)}
        continue
    if hdr is None:
        continue
    if r[0] != "":
        cur_line, cur_src = r[0], r[1].strip()
        continue
    if r[2] in ("...", "-", ""):
        continue

    def val(name):
        k = col.get(name)
        try:
            return float(r[k]) if k is not None and r[k] not in ("-", "") else 0.0
        except ValueError:
            return 0.0
    key = (cur_file, cur_line)
    a = agg.setdefault(key, {"src": cur_src, "samples": 0.0, "inst": 0.0, "tags": 0.0, "shw": 0.0, "long_sb": 0.0})
    a["samples"] += val("# Samples")
    a["inst"] += val("Instructions Executed")
    a["tags"] += val("L1 Tag Requests Global")
    a["shw"] += val("L1 Wavefronts Shared")
    a["long_sb"] += val("stall_long_sb")
tot = sum(a["samples"] for a in agg.values()) or 1.0
tinst = sum(a["inst"] for a in agg.values()) or 1.0
print("total samples %.0f, warp instructions %.3g, L1 tag requests %.3g, shared wavefronts %.3g" %
      (tot, tinst, sum(a["tags"] for a in agg.values()), sum(a["shw"] for a in agg.values())))
print("%-18s %7s %7s %9s %9s %8s  %s" % ("file:line", "samp %", "inst %", "L1 tags", "sh wavef", "long_sb", "source"))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    print("%-18s %7.1f %7.1f %9.3g %9.3g %8.0f  %s" % ("%s:%s" % (f, ln), 100 * a["samples"] / tot, 100 * a["inst"] / tinst,
                                                     a["tags"], a["shw"], a["long_sb"], a["src"][:90]))
