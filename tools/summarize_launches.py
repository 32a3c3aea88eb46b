#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (markdown table).
    python tools/summarize_launches.py launches.csv "title" [top_n]
Kernels of this library (namespace moegan) are marked; bench.py's spin kernel (the profiling lead) is left out."""
import collections, csv, re, sys


def main(path, title, top=60):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for row in r:
        full = row[ki]
        if "spin_kernel" in full:
            continue
        lib = "moegan::" in full
        name = full.split("(")[0].replace("moegan::<unnamed>::", "").replace("void ", "")
        name = re.sub(r"\b(at::native::|at::|native::|cudnn::|<unnamed>::)", "", name)[:90]
        v = float(row[vi].replace(",", ""))
        us = v / 1e3 if row[ui] in ("ns", "nsecond") else v
        a = agg.setdefault(name, [0, 0.0, lib]); a[0] += 1; a[1] += us; n += 1
    tot = sum(a[1] for a in agg.values())
    lib_n = sum(a[0] for a in agg.values() if a[2]); lib_t = sum(a[1] for a in agg.values() if a[2])
    print(f"### {title}\n")
    print(f"{n} launches, {tot:.1f} us total device time (ncu: cold cache, serialised - compare SHARES, not absolutes); library kernels "
          f"(`moegan::`): {lib_n} launches, {lib_t:.1f} us = {100 * lib_t / max(tot, 1e-9):.1f}% of the total\n")
    print("| kernel | library | launches | total us | us/launch | share |\n|---|---|---:|---:|---:|---:|")
    for name, (c, t, lib) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"| `{name}` | {'yes' if lib else ''} | {c} | {t:.1f} | {t / c:.2f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 60)
