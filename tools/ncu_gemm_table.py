#!/usr/bin/env python
"""Markdown table (one row per tcgen05 GEMM launch) from `ncu -i <rep> --page raw --csv` of an `ncu --set full` capture."""
import csv, re, sys

def main(path):
    rows = list(csv.reader(open(path)))
    hdr, unit, data = rows[0], rows[1], rows[2:]
    def col(suffix):
        return next(i for i, h in enumerate(hdr) if h.endswith(suffix))
    K, G, T = col("Kernel Name"), col("Grid Size"), col("gpu__time_duration.sum")
    R, W = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    TP = col("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")
    DT = next(i for i, h in enumerate(hdr) if h == "dram__throughput.avg.pct_of_peak_sustained_elapsed" or h.endswith(".dram__throughput.avg.pct_of_peak_sustained_elapsed"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3}
    def num(r, c):
        try:
            return float(r[c].replace(",", "")) * scale.get(unit[c], 1)
        except ValueError:
            return float("nan")
    print("| # | variant `<BN, kind, cta2, epilogue>` | grid | us | DRAM MB (r+w) | tensor-pipe active % | DRAM TB/s |\n|---:|---|---:|---:|---:|---:|---:|")
    tot_us = tot_b = 0.0
    for i, r in enumerate(data):
        m = re.search(r"umma_gemm_kernel<([^>]*)>", r[K])
        us, b = num(r, T), num(r, R) + num(r, W)
        tot_us += us; tot_b += b
        print(f"| {i} | `<{m.group(1) if m else '?'}>` | {r[G].strip('()').split(',')[0]} | {us:.1f} | {b / 1e6:.1f} | {num(r, TP):.1f} | {b / us / 1e6:.2f} |")
    print(f"\n{len(data)} launches: {tot_us:.1f} us, {tot_b / 1e6:.1f} MB DRAM traffic in total")

if __name__ == "__main__":
    main(sys.argv[1])
