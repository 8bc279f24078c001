"""Reduce an `ncu --set full` capture to the per-kernel-class numbers bench.py and DESIGN.md cite.

    ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > profiles/r2_ncu_full_raw.csv     (here, no GPU)
    python tools/ncu_traffic.py profiles/r2_ncu_full_raw.csv llama2-7b/tp1

Writes / updates profiles/r2_ncu_traffic.json:
  {"llama2-7b/tp1": {"gate_up": {"dram_bytes_per_launch": ..., "dram_read": ..., "dram_write": ...,
                                 "duration_us": ..., "dram_pct": ..., "tensor_pct": ..., "launches": n}, ...}}
Kernel class = template arguments of gemm_skinny_kernel<NT, PRO, EPI> + the grid size (qkv / o /
gate_up / down / lm_head have distinct (PRO, EPI, grid) triples), attention and the tcgen05 kernels
by name.
"""
from __future__ import annotations

import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")

WANT = {
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3,
              "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def classify(name: str, grid: float) -> str:
    m = re.search(r"gemm_skinny_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(\d+)>", name)
    if m:
        pro, epi = int(m.group(2)), int(m.group(3))
        return {(0, 0): "qkv", (1, 1): "o_or_down", (1, 2): "o_or_down_tp", (0, 3): "gate_up",
                (0, 4): "lm_head"}.get((pro, epi), f"gemm_{pro}_{epi}")
    for key, cls in (("attn_cluster_kernel", "attention"), ("prefill_gemm_tc_kernel", "prefill_tc"),
                     ("lmhead_tc_kernel", "lmhead_tc"), ("tp_allreduce_ll_kernel", "allreduce_ll"),
                     ("rms_canon_kernel", "rms_canon")):
        if key in name:
            return cls
    return "other"


def main():
    path, key = sys.argv[1], sys.argv[2]
    rows = {}
    with open(path, newline="") as f:
        table = [r for r in csv.reader(l for l in f if not l.startswith("=="))]
    header, units = table[0], table[1]          # `--page raw --csv`: one column per metric, a units row
    col = {h: i for i, h in enumerate(header)}
    for r in table[2:]:
        ent = {"name": r[col["Kernel Name"]]}
        g = re.findall(r"\d+", r[col["Grid Size"]])
        ent["grid"] = float(g[0]) * float(g[1]) * float(g[2]) if len(g) == 3 else 0.0
        for metric, short in WANT.items():
            if metric not in col or short == "grid":
                continue
            try:
                val = float(r[col[metric]].replace(",", ""))
            except ValueError:
                continue
            ent[short] = val * UNIT_SCALE.get(units[col[metric]], 1.0)
        rows[r[col["ID"]]] = ent
    agg = {}
    for ent in rows.values():
        cls = classify(ent["name"], ent.get("grid", 0))
        if cls == "o_or_down":      # same instantiation: tell them apart by DRAM bytes (33.6 vs 90.2 MB at 7B)
            cls = "down" if ent.get("dram_read", 0) > 60e6 else "o_proj"
        a = agg.setdefault(cls, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "duration_us": 0.0,
                                 "dram_pct": 0.0, "tensor_pct": 0.0, "warps_active_pct": 0.0, "regs": 0})
        a["launches"] += 1
        a["dram_read"] += ent.get("dram_read", 0.0)
        a["dram_write"] += ent.get("dram_write", 0.0)
        a["duration_us"] += ent.get("duration", 0.0)
        a["dram_pct"] += ent.get("dram_pct", 0.0)
        a["tensor_pct"] += ent.get("tensor_pct", 0.0)
        a["warps_active_pct"] += ent.get("warps_active_pct", 0.0)
        a["regs"] = int(ent.get("regs", 0))
    for a in agg.values():
        n = a["launches"]
        for k in ("dram_read", "dram_write", "duration_us", "dram_pct", "tensor_pct", "warps_active_pct"):
            a[k] /= n
        a["dram_bytes_per_launch"] = a["dram_read"] + a["dram_write"]
    tab = {}
    if os.path.exists(OUT):
        tab = json.load(open(OUT))
    tab[key] = agg
    tab.setdefault("_source", {})[key] = os.path.relpath(path, ROOT)
    with open(OUT, "w") as f:
        json.dump(tab, f, indent=1, sort_keys=True)
    for cls, a in sorted(agg.items()):
        print(f"{cls:12s} n={a['launches']:3d} dur={a['duration_us']:8.2f} us dram={a['dram_bytes_per_launch'] / 1e6:8.2f} MB "
              f"dram%={a['dram_pct']:5.1f} tensor%={a['tensor_pct']:5.1f} warps%={a['warps_active_pct']:5.1f} regs={a['regs']}")


if __name__ == "__main__":
    main()
