#!/usr/bin/env python
"""Condense `ncu -i <report>.ncu-rep --page raw --csv` into the few counters the roofline needs, per kernel launch.

usage: ncu_summary.py raw.csv [out.json] [--workload n=10000000,rank=32,n_gpus=1] [--source "text"]
The JSON has the layout bench.py reads (profiles/r2_ncu_traffic.json): kernels keyed by the library's kernel-class name.
"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}
CLASS = [(r"k_mc_step", "k_mc_step"), (r"k_mc_spmm", "k_mc_spmm"), (r"k_mc_combine", "k_mc_dir"), (r"k_put_rows", "k_exchange"),
         (r"k_spmm<", "k_spmm"), (r"k_uvt", "k_uvt"), (r"k_con_gather", "k_gather"), (r"k_wsum", "k_wsum"),
         (r"k_dense_", "k_dense_dmma"), (r"k_gram_dmma", "k_dense_dmma"), (r"k_reduce", "k_reduce"), (r"k_map", "k_vec")]
WANT = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes", "gpu__time_duration.sum": "time_ms",
        "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "launch__occupancy_limit_registers": "occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_inst_pct",
        "launch__shared_mem_per_block_dynamic": "dynamic_smem", "launch__grid_size": "grid", "launch__block_size": "block",
        "smsp__cycles_active.avg": "smsp_cycles_active", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct"}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = dict(a[2:].split("=", 1) for a in sys.argv[1:] if a.startswith("--") and "=" in a)
    rows = list(csv.reader(open(args[0])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kernels, launches = {}, []
    for r in data:
        name = r[col["Kernel Name"]]
        ent = {"kernel": name[:70]}
        for metric, key in WANT.items():
            if metric in col and r[col[metric]] not in ("", "n/a"):
                try:
                    v = float(r[col[metric]].replace(",", ""))
                except ValueError:
                    continue
                u = units[col[metric]]
                if key.endswith("_bytes") or key == "time_ms":
                    v *= UNIT.get(u, 1.0)
                ent[key] = v
        cls = next((c for pat, c in CLASS if re.search(pat, name)), None)
        ent["class"] = cls
        launches.append(ent)
        if cls and cls not in kernels:
            kernels[cls] = ent
    wl = {}
    for kv in opts.get("workload", "").split(","):
        if "=" in kv:
            k, v = kv.split("=")
            wl[k] = int(v)
    out = {"source": opts.get("source", f"ncu --set full --clock-control none; raw page {args[0]}"), "workload": wl, "kernels": kernels,
           "launches": launches}
    text = json.dumps(out, indent=1)
    if len(args) > 1:
        open(args[1], "w").write(text + "\n")
    for e in launches:
        dr, dw, t = e.get("dram_read_bytes", 0), e.get("dram_write_bytes", 0), e.get("time_ms", 0)
        gbs = (dr + dw) / (t * 1e-3) / 1e9 if t else 0
        print(f"{e['kernel'][:48]:48s} {t:8.3f} ms  rd {dr / 1e9:7.2f} GB  wr {dw / 1e9:7.2f} GB  {gbs:7.0f} GB/s  regs {e.get('registers', 0):.0f}  "
              f"warps {e.get('warps_active_pct', 0):.1f}%  L2hit {e.get('l2_hit_pct', 0):.1f}%")


if __name__ == "__main__":
    main()
