#!/usr/bin/env python
"""Per-kernel SASS evidence of the built library: which kernels carry tcgen05 (UTCHMMA), TMEM loads (LDTM), TMA loads / stores (UTMALDG /
UTMASTG), multicast commits (UTCBAR), packed fp32 FMA (FFMA2) - counted with cuobjdump, no GPU needed.

    python tools/sass_summary.py > profiles/sass_summary.md
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "moe-gan_cpsc541_b200", "libmoegan_sm100.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "FFMA2", "FFMA", "HFMA2", "MUFU", "MEMBAR.ALL.GPU", "CCTL.IVALL"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); per[cur] = collections.Counter(); continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                per[cur][k] += 1
    dm = demangle(list(per))
    print("# SASS summary of libmoegan_sm100.so (cuobjdump -sass; sm_100a)\n")
    print("Instruction mnemonics that prove the Blackwell paths: `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld (TMEM -> registers), "
          "`UTMALDG` / `UTMASTG` = TMA tensor load / store, `UTCBAR` = tcgen05.commit (mbarrier arrive, `.MULTICAST` for CTA pairs), `SYNCS` = mbarrier ops, "
          "`FFMA2` = packed fp32 FMA.\n")
    print("| kernel | instr | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    tot = collections.Counter()
    for fn, c in per.items():
        name = re.sub(r"moegan::\(anonymous namespace\)::", "", dm.get(fn, fn))
        name = re.sub(r"\(.*", "", name)
        if not any(c[k] for k in KEYS[:7] + ["FFMA2"]):
            continue
        print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
        tot.update(c)
    print("\nTotals over the listed kernels: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
