"""SASS evidence for the in-tree library: per-kernel counts of the mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit ->
UTCBAR, legacy mma.sync -> HMMA) plus the global atomics (REDG/ATOMG) and the register / shared-memory footprint.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt        (no GPU needed: cuobjdump on the .so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "medicalseg_b200", "lib", "libmedseg_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "HMMA", "REDG", "ATOMG",
        "SYNCS", "LDGSTS", "STG", "LDG"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    counts, order = collections.defaultdict(collections.Counter), []
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            order.append(name)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            op = m.group(1)
            counts[name]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[name][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass %s  (arch sm_100a)" % os.path.relpath(lib, ROOT))
    print("# per kernel: instruction count, registers, static shared bytes, then the mnemonic counts that are non-zero")
    tot = collections.Counter()
    for mangled, nice in zip(order, demangled):
        c = counts[mangled]
        nice = re.sub(r"\(.*$", "", nice)
        reg, sh = usage.get(mangled, (0, 0))
        marks = " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])
        print("%-96s inst=%-6d reg=%-3d smem=%-6d %s" % (nice[:96], c["_total"], reg, sh, marks))
        tot.update({k: c[k] for k in KEYS})
    print("# TOTAL " + " ".join("%s=%d" % (k, tot[k]) for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
