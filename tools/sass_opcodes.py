"""profiles/r02_sass_opcodes.txt: per-kernel counts of the Blackwell-specific SASS opcodes in the built library
(cuobjdump -sass pyvb_b200/libpyvb_b200.so): UTCIMMA / UTCHMMA (tcgen05.mma kind::i8 / kind::f16, .2CTA = cta_group::2),
LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk copy), DMMA (FP64 tensor core), SYNCS (mbarrier).
usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyvb_b200", "libpyvb_b200.so")
OPS = ["UTCIMMA.2CTA", "UTCIMMA", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "DFMA", "SYNCS", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::|pyvb::", "", fn).split("(")[0].replace("void ", "")
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        total[fn] += 1
        for key in OPS:
            if op == key or op.startswith(key + "."):
                if key == "UTCIMMA" and ".2CTA" in op:
                    continue
                counts[fn][key] += 1
                break
        if op.startswith("UTCIMMA") and ".2CTA" in op:
            counts[fn]["UTCIMMA.2CTA"] += 1
    print("SASS opcode counts per kernel, %s (sm_100a), cuobjdump -sass" % os.path.relpath(LIB, ROOT))
    print("%-72s %7s  %s" % ("kernel", "instr", "  ".join("%s" % k for k in OPS)))
    tot = collections.Counter()
    for fn, c in counts.items():
        if not any(c[k] for k in OPS if k not in ("DFMA", "LDGSTS")):
            continue
        print("%-72s %7d  %s" % (fn[:72], total[fn], "  ".join("%*d" % (len(k), c[k]) for k in OPS)))
        tot.update(c)
    print("%-72s %7s  %s" % ("ALL KERNELS", "", "  ".join("%*d" % (len(k), tot[k]) for k in OPS)))


if __name__ == "__main__":
    main()
