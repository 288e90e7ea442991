"""Compact per-launch table from `ncu --set full ... --page raw --csv` exports (tools/collect_profiles*.sh):
time, DRAM bytes and GB/s (and % of the measured HBM peak), tensor-pipe and tensor-memory activity, issue activity,
registers, achieved occupancy.  usage: python tools/ncu_summary.py gpurun_out/r2_ncu_*.csv > profiles/r2_ncu_summary.txt"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    HBM = 6650.0

COLS = [("gpu__time_duration.sum", "us", 1.0), ("dram__bytes_read.sum", "rdMB", 1.0), ("dram__bytes_write.sum", "wrMB", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1.0),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tmem%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("lts__t_sector_hit_rate.pct", "L2hit%", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1.0),
        ("launch__grid_size", "grid", 1.0)]


def short(name):
    name = re.sub(r"\(bool\)|\(int\)|escb::|tc::|mf::", "", name)
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0][:78]


def to_unit(v, unit, want):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
             "nsecond": 1e-3, "msecond": 1e3}
    return v * scale.get(unit, 1.0)


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path, errors="replace")))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"== {os.path.basename(path)}   (peak HBM {HBM:.0f} GB/s, MEASURED_PEAKS.json)")
    print(f"{'kernel':80s} " + " ".join(f"{n:>8s}" for _, n, _ in COLS) + "   GB/s  %HBM")
    for r in rows[2:]:
        vals = []
        for key, _, _ in COLS:
            vals.append(to_unit(r[ix[key]], units[ix[key]], None) if key in ix else float("nan"))
        us, rd, wr = vals[0], vals[1], vals[2]
        gbs = (rd + wr) * 1e6 / (us * 1e-6) / 1e9 if us > 0 else 0.0
        print(f"{short(r[ix['Kernel Name']]):80s} " + " ".join(f"{v:8.1f}" for v in vals) + f" {gbs:6.0f} {100 * gbs / HBM:5.1f}")
    print()
