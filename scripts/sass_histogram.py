"""SASS opcode histogram of the tensor-core kernels in the built library (cuobjdump -sass | c++filt), the evidence that the
fused kernels carry tcgen05 / TMEM / mbarrier instructions:  python scripts/sass_histogram.py > profiles/r02_sass_mlp_tc.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "indoor-nerf_b200", "csrc", "libpocketnerf.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "SYNCS", "USETMAXREG", "ELECT", "UBLKPF", "REDG", "ATOMG", "SHFL", "LDG", "STG",
       "LDS", "STS", "LDL", "STL", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\s+Function : \S+\n", sass)[1:]
    print("# SASS opcode histogram of the tensor-core kernels in libpocketnerf.so (cuobjdump -sass, sm_100a), scripts/sass_histogram.py.")
    print("# UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier,")
    print("# USETMAXREG = setmaxnreg, ELECT = elect.sync, UBLKPF = cp.async.bulk.prefetch.L2, REDG = red.global, ATOMG = atom.global,")
    print("# LDL / STL = local memory (spills)\n")
    for name, body in zip(names, blocks):
        if not re.search(r"mlp_tc_|field_bwd4|tc_selftest", name):
            continue
        ops = collections.Counter()
        for line in body.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        print(name)
        print("  %d instructions; tensor/TMEM/async: %s" % (total, ", ".join("%s %d" % (k, ops[k]) for k in KEY if ops[k])))
        print("  top: " + ", ".join("%s %d" % kv for kv in ops.most_common(12)) + "\n")


if __name__ == "__main__":
    sys.exit(main())
