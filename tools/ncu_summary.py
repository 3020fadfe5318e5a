#!/usr/bin/env python
"""Summarise ncu output for profiles/: either a launch list (--metrics gpu__time_duration.sum --csv log) grouped per
kernel with shares, or a `ncu -i X.ncu-rep --page raw --csv` dump reduced to the columns the roofline needs.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.csv
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_summary.py full - > profiles/rNN_full.csv
"""
import csv
import sys
from collections import OrderedDict

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cycles_pct"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_inst_pct"),
        ("sm__cycles_active.avg", "sm_cycles_active"), ("sm__cycles_elapsed.avg", "sm_cycles_elapsed"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def to_us(v, unit):
    m = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}
    return float(v) * m.get(unit, 1)


def read_rows(path):
    f = sys.stdin if path == "-" else open(path)
    rows = [r for r in csv.reader(l for l in f if not l.startswith("==")) if r]
    return rows


def launches(path):
    rows = read_rows(path)
    h = rows[0]
    ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0]
        us = to_us(r[iv].replace(",", ""), r[iu])
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,avg_us,max_us,share_pct")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name},{a[0]},{a[1]:.1f},{a[1] / a[0]:.1f},{a[2]:.1f},{100 * a[1] / tot:.1f}")


def full(path):
    rows = read_rows(path)
    h, u = rows[0], rows[1]
    ik = h.index("Kernel Name")
    cols = [(h.index(c), n, c) for c, n in COLS if c in h]
    print("kernel," + ",".join(n + ("_us" if n == "time" else "_MB" if n.startswith("dram_") and not n.endswith("pct") else "")
                               for _, n, _ in cols))
    for r in rows[2:]:
        out = [r[ik].split("(")[0]]
        for i, n, c in cols:
            v = r[i].replace(",", "")
            if n == "time":
                out.append(f"{to_us(v, u[i]):.1f}")
            elif n in ("dram_rd", "dram_wr"):
                out.append(f"{to_bytes(v, u[i]) / 1e6:.1f}")
            else:
                try:
                    out.append(f"{float(v):.1f}")
                except ValueError:
                    out.append(v)
        print(",".join(out))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
