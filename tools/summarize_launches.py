#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (markdown table)."""
import collections, csv, sys

def main(path, title):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    n = 0
    for row in r:
        name = row[ki].split("(")[0].replace("moegan::<unnamed>::", "").replace("void ", "")[:80]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(row[vi].replace(",", "")); n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"### {title}\n\n{n} launches, {tot / 1e3:.1f} us total device time (ncu: cold cache, serialised - compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|")
    for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print(f"| `{name}` | {c} | {t / 1e3:.1f} | {t / c / 1e3:.2f} | {100 * t / tot:.1f}% |")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
