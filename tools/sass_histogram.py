"""SASS opcode histogram of libsva_b200.so (cuobjdump, no GPU needed): which Blackwell-specific instructions the library actually contains.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stereovisionarray_b200", "libsva_b200.so")
WATCH = ["VIADDMNMX", "VIMNMX3", "VIMNMX", "CREDUX", "VABSDIFF4", "IDP", "POPC", "REDG", "RED", "ATOMG", "LDGSTS", "UBLKCP", "UBLKRED", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS",
         "TEX", "TLD", "SHFL", "PRMT", "LDS", "STS", "LDG", "STG", "IMAD", "LEA", "IADD3", "LOP3", "FENCE", "MEMBAR", "BAR", "ELECT", "R2UR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per_kernel = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per_kernel.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    total = collections.Counter()
    for c in per_kernel.values():
        total.update(c)
    print("# SASS opcode histogram of stereovisionarray_b200/libsva_b200.so (sm_100a), %d kernels, %d instructions" % (len(per_kernel), sum(total.values())))
    print("# watched opcodes (whole library):")
    for op in WATCH:
        print("%-14s %8d" % (op, total.get(op, 0)))
    print("\n# kernels that contain the copy-engine / mbarrier / DPX / texture opcodes (demangle with c++filt):")
    for k, c in per_kernel.items():
        tags = {op: c[op] for op in ("UBLKCP", "UBLKRED", "SYNCS", "VIADDMNMX", "CREDUX", "VABSDIFF4", "POPC", "TEX", "LDGSTS", "REDG") if c.get(op)}
        if tags:
            print("%-90s %s" % (k[:90], " ".join("%s=%d" % kv for kv in tags.items())))
    print("\n# top 40 opcodes:")
    for op, n in total.most_common(40):
        print("%-14s %8d" % (op, n))


if __name__ == "__main__":
    main()
