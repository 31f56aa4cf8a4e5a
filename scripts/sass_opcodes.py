#!/usr/bin/env python
"""Counts the SASS opcodes that prove (or disprove) a Blackwell-native kernel set in libfemasr_b200.so
(B200_PROFILING.md "What proves a Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
TMA -> UTMALDG/UTMASTG/UBLKCP, legacy mma.sync -> HMMA.  Runs without a GPU (cuobjdump on the in-tree library).

    python scripts/sass_opcodes.py > profiles/sass_opcodes_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "femasr_b200", "libfemasr_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|UTCCP|UTMALDG(?:\.[0-9]D)?(?:\.2CTA)?(?:\.MULTICAST)?|"
                 r"UTMASTG(?:\.[0-9]D)?|UTMAREDG|UBLKCP|LDTM|STTM|HMMA\.[0-9]+|LDSM|SYNCS(?:\.[A-Z0-9_]+)*|UCGABAR_[A-Z]+|"
                 r"MUFU\.[A-Z0-9]+|FFMA|LDG|STG|LDS|STS|ATOMG|RED)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per_kernel = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per_kernel.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        for op in PAT.findall(line):
            cur[op] += 1
    total = collections.Counter()
    for c in per_kernel.values():
        total.update(c)
    keys = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "UTMALDG.4D", "UTMALDG.2D.2CTA", "UTMALDG.4D.2CTA", "UTMASTG.2D",
            "UTMASTG.4D", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "HMMA.16816", "LDSM"]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(per_kernel)} kernels)")
    print("## whole library")
    for k in keys:
        print(f"{k:28s} {total.get(k, 0)}")
    other = sorted(k for k in total if k.startswith(("UTC", "UTMA")) and k not in keys)
    for k in other:
        print(f"{k:28s} {total[k]}")
    print("## kernels using the tensor cores / TMA / TMEM")
    for name, c in per_kernel.items():
        tc = {k: v for k, v in c.items() if k.startswith(("UTC", "UTMA", "UBLKCP", "LDTM", "STTM", "HMMA", "LDSM"))}
        if tc:
            short = name if len(name) < 150 else name[:147] + "..."
            print(short)
            print("    " + ", ".join(f"{k} {v}" for k, v in sorted(tc.items())))


if __name__ == "__main__":
    sys.exit(main())
