#!/usr/bin/env python3
"""Per-kernel counts of the Blackwell-specific SASS instructions in libbclinear.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (1-D bulk copy), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), FFMA2 (packed fp32 FMA, sm_100), plus FFMA as the CUDA-core reference.  Writes a table to stdout; `> profiles/rN_sass_summary.txt`."""
import collections
import os
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama-fft_b200", "libbclinear.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTMAPF", "SYNCS", "FFMA2", "FFMA", "REDG", "LDG", "STG"]   # (FFMA2 = packed fma.rn.f32x2, before its prefix FFMA)


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
                    break
            counts[cur]["_total"] += 1
    dm = demangle(list(counts))
    print("# cuobjdump -sass %s  (sm_100a); instruction counts per kernel" % os.path.basename(LIB))
    print("%-110s %7s " % ("kernel", "instrs") + " ".join("%8s" % o for o in OPS))
    tot = collections.Counter()
    for k, c in counts.items():
        name = re.sub(r"\(.*", "", dm.get(k, k))
        name = name.replace("void ", "")
        print("%-110s %7d " % (name[:110], c["_total"]) + " ".join("%8d" % c[o] for o in OPS))
        tot.update(c)
    print("%-110s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[o] for o in OPS))


if __name__ == "__main__":
    main()
