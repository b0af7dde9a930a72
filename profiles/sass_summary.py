#!/usr/bin/env python
"""Static view of the built library: per kernel of xz_b200/libxzb200.so the register / stack / shared-memory use
(cuobjdump -res-usage) and the SASS instruction mix (cuobjdump -sass): which memory, warp-collective and
synchronisation instructions the hand-written kernels really contain.  Runs without a GPU.

    python profiles/sass_summary.py > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "xz_b200", "libxzb200.so")
GROUPS = [
    ("global load", r"^LDG"), ("  of which 128-bit", r"^LDG\S*\.128"), ("global store", r"^STG"), ("  of which 128-bit", r"^STG\S*\.128"),
    ("shared load", r"^LDS"), ("  of which 128-bit", r"^LDS\S*\.128"), ("shared store", r"^STS"), ("  of which 128-bit", r"^STS\S*\.128"),
    ("local (stack) ld/st", r"^(LDL|STL)"), ("global atomics / reductions", r"^(ATOMG|RED|ATOM)\b|^(ATOMG|RED)\."), ("shared atomics", r"^ATOMS"),
    ("warp shuffle", r"^SHFL"), ("warp vote / match / redux", r"^(VOTE|VOTEU|MATCH|REDUX)"), ("warp barrier", r"^(WARPSYNC|BSYNC|BSSY)"),
    ("CTA barrier", r"^BAR"), ("memory fence", r"^(MEMBAR|FENCE)"), ("ld.acquire / st.release style (.STRONG)", r"^(LDG|STG|LDS|STS|LD|ST)\S*\.STRONG"),
    ("prefetch (CCTL / LDG to RZ)", r"^CCTL"), ("bulk copy / TMA (UBLKCP, UTMALDG, UTMASTG)", r"^(UBLKCP|UTMA)"), ("mbarrier (SYNCS)", r"^SYNCS"),
    ("tensor core (UTC*MMA, HMMA, IMMA)", r"^(UTC|HMMA|IMMA|QGMMA|HGMMA)"), ("branches", r"^(BRA|BRX|JMP|CALL|RET)"),
    ("integer multiply-add (IMAD)", r"^IMAD"), ("3-input logic (LOP3)", r"^LOP3"), ("funnel shift / shift", r"^(SHF|SHL|SHR)"), ("find-leading / popc / brev", r"^(FLO|POPC|BREV)"),
]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in ln:
            usage[cur] = ln.strip()
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur:
            per[cur].append(m.group(1))
    print("# SASS summary of xz_b200/libxzb200.so (sm_100a), hand-written kernels only (CUB's sort / select kernels left out)")
    print("# produced by profiles/sass_summary.py; counts are static instructions, not executed ones\n")
    for fn, ins in per.items():
        if "xzb_k_" not in fn:
            continue
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"== {dem}")
        print(f"   {usage.get(fn, '')}")
        print(f"   instructions: {len(ins)}")
        for name, pat in GROUPS:
            n = sum(1 for i in ins if re.search(pat, i))
            if n:
                print(f"   {name:<48}{n}")
        print()


if __name__ == "__main__":
    main()
