"""SASS opcode histogram of the built library, per kernel (cuobjdump -sass): the evidence for which tensor / copy
paths the shipped kernels use.   python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gp-plus_b200", "gpplus_b200", "lib", "libgpplus_b200.so")
COLS = ["DMMA", "UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "UTCBAR", "SYNCS", "DFMA",
        "DMUL", "DADD", "I2F", "MUFU", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "ATOM", "RED"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per, order, cur, k = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[k]) if k < len(names) else m.group(1)
            k += 1
            per[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    print("SASS opcode histogram of %s (cuobjdump -sass, sm_100a), per kernel" % os.path.relpath(lib, ROOT))
    print("columns: total instructions | " + " ".join(COLS))
    tot = collections.Counter()
    for name in sorted(order, key=lambda n: -per[n]["_total"]):
        c = per[name]
        tot.update(c)
        print("%-90s %6d | %s" % (name[:90], c["_total"], " ".join("%s=%d" % (o, c[o]) for o in COLS if c[o])))
    print("%-90s %6d | %s" % ("ALL KERNELS", tot["_total"], " ".join("%s=%d" % (o, tot[o]) for o in COLS)))


if __name__ == "__main__":
    main()
