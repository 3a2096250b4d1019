"""Counts the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,
TMA -> UTMALDG, FP64 tensor pipe -> DMMA) per kernel of cosma_b200/lib/libcosma_b200.so. No GPU needed.
    python tools/sass_evidence.py > profiles/<round>_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cosma_b200", "lib", "libcosma_b200.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCMMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|DMMA|HMMA|HGMMA|LDGSTS|SYNCS)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, kern = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        if kern:
            counts[kern].update(PAT.findall(line))
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonics per kernel of libcosma_b200.so (sm_100a), from `cuobjdump -sass`; kernels without any of them omitted")
    for (k, c), name in zip(counts.items(), names):
        if c:
            short = re.sub(r"\(CUtensorMap_st.*", "(...)", name)
            print("%-110s %s" % (short[:110], " ".join("%s=%d" % kv for kv in sorted(c.items()))))
    return 0


if __name__ == "__main__":
    sys.exit(main())
