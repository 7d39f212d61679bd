#!/usr/bin/env python
"""SASS instruction histogram of the shipped library, per kernel family: proof on the page that the contraction path is tcgen05 /
TMEM / TMA (UTCHMMA, LDTM, UTMALDG, UTMASTG) — `cuobjdump -sass` of every object linked into indm_b200/libindm_b200.so.
    python tools/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "indm_b200", "csrc")
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "ELECT", "HMMA", "IMMA", "FFMA", "MUFU", "SHFL",
       "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "BAR"]


def main():
    objs = sorted(f for f in os.listdir(CSRC) if f.endswith(".o"))
    fam = collections.OrderedDict()
    for o in objs:
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, o)], capture_output=True, text=True).stdout
        cur = None
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(anonymous namespace\)::", "", name)
                name = re.sub(r"^void ", "", name).split("(")[0]
                base = re.sub(r"<.*", "", name)
                cur = fam.setdefault((o, base), dict(n=0, ops=collections.Counter(), total=0))
                cur["n"] += 1
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur is not None:
                op = m.group(1)
                cur["total"] += 1
                root = op.split(".")[0]
                cur["ops"][root] += 1
                if op.startswith("UTCHMMA.2CTA") or ".2CTA" in op and root == "UTCHMMA":
                    cur["ops"]["UTCHMMA.2CTA"] += 1
    print("# SASS instruction histogram of libindm_b200.so (sm_100a), per kernel family\n")
    print("`cuobjdump -sass` of every object in `indm_b200/csrc/`, instruction roots counted over ALL template instantiations of a kernel "
          "(`n` = number of instantiations).  UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), LDTM = tcgen05.ld (TMEM -> registers), "
          "UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  Made by `tools/sass_histogram.py`.\n")
    print("| object | kernel | n | instr | " + " | ".join(KEY) + " |")
    print("|---|---|---|---|" + "---|" * len(KEY))
    tot = collections.Counter()
    for (o, base), d in fam.items():
        print(f"| {o} | `{base}` | {d['n']} | {d['total']} | " + " | ".join(str(d['ops'].get(k, 0)) for k in KEY) + " |")
        for k in KEY:
            tot[k] += d["ops"].get(k, 0)
    print("| **all** | | | | " + " | ".join(str(tot[k]) for k in KEY) + " |")


if __name__ == "__main__":
    main()
