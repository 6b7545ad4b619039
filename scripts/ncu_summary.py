#!/usr/bin/env python
"""Summarise `ncu --set full` captures (.ncu-rep) into a markdown table + the two JSON files bench.py reads.

usage: python scripts/ncu_summary.py OUT.md REP [REP ...] [--json]
Per kernel name the LAST captured launch is reported (later launches are warmer).  With --json the
per-launch DRAM bytes and the headline metrics are merged into profiles/ncu_traffic.json and
profiles/ncu_metrics.json under the short kernel names bench.py uses.
"""
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PEAK_GBS = 6555.5  # MEASURED_PEAKS.json hbm_gbs on this pool

COLS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0),
    ("FMA pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    ("L1 throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L1 LSU wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct", 1.0),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
    ("long_sb / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
    ("warp inst", "smsp__inst_executed.sum", 1.0),
    ("DRAM rd MB", "dram__bytes_read.sum", 1.0),
    ("DRAM wr MB", "dram__bytes_write.sum", 1.0),
    ("dyn smem KB", "launch__shared_mem_per_block_dynamic", 1.0),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
        "nsecond": 1e-3, "Kbyte/block": 1e3, "byte/block": 1.0}


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    s = m.group(1) if m else name
    return {"k_desc_radial2": "k_desc_radial", "k_force_final2": "k_force_final"}.get(s, s)


def load(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(head, units, r):
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                d[h] = v
                continue
            d[h] = x * UNIT.get(u, 1.0) if u in UNIT else x
        res[d["Kernel Name"]] = d  # last launch of each name wins
    return res


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out_md, reps = Path(args[0]), args[1:]
    kernels = {}
    for rep in reps:
        for k, d in load(rep).items():
            kernels[k] = (Path(rep).name, d)
    lines = ["| kernel | capture | " + " | ".join(c[0] for c in COLS) + " | DRAM GB/s (of measured %.1f) |" % PEAK_GBS,
             "|" + "---|" * (len(COLS) + 3)]
    traffic, metrics = {}, {}
    for k, (rep, d) in kernels.items():
        cells = []
        for label, key, _ in COLS:
            v = d.get(key)
            if v is None or isinstance(v, str):
                cells.append("-")
            elif label.endswith("MB"):
                cells.append("%.0f" % (v / 1e6))
            elif label.endswith("KB"):
                cells.append("%.1f" % (v / 1e3))
            elif label == "warp inst":
                cells.append("%.3g" % v)
            elif label in ("us", "regs"):
                cells.append("%.0f" % v)
            else:
                cells.append("%.1f" % v)
        tot = (d.get("dram__bytes_read.sum") or 0.0) + (d.get("dram__bytes_write.sum") or 0.0)
        us = d.get("gpu__time_duration.sum") or 1.0
        gbs = tot / us / 1e3
        name = re.sub(r"^void\s+|\(anonymous namespace\)::|unnamed>::", "", k)
        lines.append("| `%s` | %s | " % (name.split("(")[0], rep) + " | ".join(cells) + " | %.0f (%.0f %%) |" % (gbs, 100 * gbs / PEAK_GBS))
        s = short(k)
        traffic[s] = tot
        metrics[s] = {
            "issue_slots_busy_pct": d.get(COLS[3][1]), "fma_pipe_pct": d.get(COLS[4][1]),
            "tensor_pipe_active_pct": d.get(COLS[5][1]), "l1_lsu_wavefronts_pct": d.get(COLS[7][1]),
            "l2_hit_pct": d.get(COLS[9][1]), "l1_hit_pct": d.get(COLS[8][1]),
            "dram_read_MB": (d.get("dram__bytes_read.sum") or 0.0) / 1e6,
            "dram_write_MB": (d.get("dram__bytes_write.sum") or 0.0) / 1e6,
            "achieved_occupancy_pct": d.get(COLS[2][1]), "registers": d.get(COLS[1][1]),
            "warp_instructions": d.get("smsp__inst_executed.sum"),
            "stall_long_scoreboard_per_issue": d.get(COLS[10][1]), "ncu_us": us, "capture": rep,
        }
    out_md.write_text("\n".join(lines) + "\n")
    print("\n".join(lines))
    if "--json" in sys.argv:
        tf, mf = ROOT / "profiles" / "ncu_traffic.json", ROOT / "profiles" / "ncu_metrics.json"
        t = json.loads(tf.read_text())
        m = json.loads(mf.read_text())
        t.update(traffic)
        m.update(metrics)
        if "k_mlp_tc" in metrics:
            m["k_mlp"] = dict(metrics["k_mlp_tc"], kernel="k_mlp_tc")
            t["k_mlp"] = traffic["k_mlp_tc"]
        src = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, 1 000 000-atom PbTe; " + ", ".join(Path(r).name for r in reps)
        t["_source"] = src
        m["_source"] = "scripts/ncu_summary.py over " + ", ".join(Path(r).name for r in reps)
        tf.write_text(json.dumps(t, indent=1) + "\n")
        mf.write_text(json.dumps(m, indent=1) + "\n")


if __name__ == "__main__":
    main()
