#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics
the roofline discussion needs.  usage: python profiles/ncu_summary.py <file.ncu-rep> [regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not pat.search(name):
            continue
        print(f"== {name[:90]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
            if any(h.endswith(k) or short == k for k in KEYS) or "warp_issue_stalled" in h and \
                    h.endswith("_per_warp_active.pct"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if "warp_issue_stalled" in h and v < 2.0:
                    continue
                print(f"   {short:75s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
