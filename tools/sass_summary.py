"""Per-kernel SASS evidence of the Blackwell-native paths in libemu_b200.so: counts of UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store), UBLKCP, HMMA (mma.sync), FFMA2 / FADD2 (packed fp32), USETMAXREG.
Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt   (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "emu_b200", "libemu_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "FFMA2", "FADD2", "USETMAXREG", "MUFU.EX2",
       "STS.128", "LDS.128", "SYNCS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts, order = None, collections.defaultdict(collections.Counter), []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kern = kern.replace("(anonymous namespace)::", "").replace("emu::", "")
            kern = re.sub(r"\(.*", "", kern)
            order.append(kern)
            continue
        if kern is None:
            continue
        for p in PAT:
            if re.search(r"\b" + re.escape(p), line):
                counts[kern][p] += 1
    print("# SASS mnemonic counts per kernel of emu_b200/libemu_b200.so (sm_100a); kernels without any of the mnemonics omitted")
    print("%-72s %s" % ("kernel", " ".join("%9s" % p for p in PAT)))
    for k in order:
        c = counts[k]
        if not any(c.values()):
            continue
        print("%-72s %s" % (k[:72], " ".join("%9d" % c[p] for p in PAT)))


if __name__ == "__main__":
    main()
