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


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[1] == "--stalls"):
    main()


def stalls(rep, top=25):
    """warp-stall sampling totals and the hottest CUDA source lines (needs -lineinfo +
    --import-source on): python profiles/ncu_summary.py --stalls <file.ncu-rep>"""
    import collections

    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                          "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernel, hdr, cur = None, None, None
    agg, st = collections.Counter(), collections.Counter()
    for r in rows:
        if r and r[0] == "Kernel Name":
            if agg:
                break           # first kernel of the report only
            kernel = r[1]
            continue
        if "# Samples" in r:
            hdr = r
            i_n = hdr.index("# Samples")
            cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "" and r[2] == "-":
            cur = (r[0], r[1].strip()[:100])
        elif r[0] == "" and r[2].startswith("0x"):
            try:
                agg[cur] += int(r[i_n])
                for i in cols:
                    st[hdr[i]] += int(r[i])
            except ValueError:
                pass
    tot = max(1, sum(agg.values()))
    print(f"# kernel: {kernel}")
    print(f"# warp stall sampling, {tot} samples")
    for k, v in st.most_common(10):
        print(f"   {k:28s} {v:9d} {100.0 * v / tot:5.1f} %")
    print("# hottest source lines (share of all samples)")
    for (ln, src), n in agg.most_common(top):
        print(f"   {100.0 * n / tot:5.2f} %  line {ln:>5s}  {src}")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "--stalls":
    stalls(sys.argv[2])
