"""Per-kernel count of the SASS mnemonics that prove the hardware path (tcgen05 MMA / TMEM loads / TMA loads and stores /
bulk copies / legacy MMA) in the built library: `python tools/sass_summary.py > profiles/rNN_sass.txt`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "modelcompose_b200", "_lib", "libmodelcompose_b200.so")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCQMMA|UTCMMA|LDTM(?:\.[A-Za-z0-9x]+)*|STTM(?:\.[A-Za-z0-9x]+)*|UTMALDG(?:\.[0-9]D)?|UTMASTG(?:\.[0-9]D)?|"
                 r"UTMAPF|UBLKCP(?:\.[A-Z]+)*|UBLKPF|UTCBAR(?:\.[A-Z0-9]+)*|UTCCP|HMMA\.[0-9]+\.F32(?:\.BF16)?|LDSM(?:\.[0-9A-Z]+)*|"
                 r"SYNCS(?:\.[A-Z_]+)*|LDG\.E(?:\.[A-Z0-9]+)*\.128(?:\.[A-Z]+)*|STG\.E(?:\.[A-Z0-9]+)*\.128|MUFU\.EX2|ELECT)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for tok in PAT.findall(line):
            kernels[cur][tok] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)")
    print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor load/store,")
    print("# UBLKCP = cp.async.bulk (1-D bulk copy), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync, LDSM = ldmatrix")
    for (name, cnt), dm in zip(kernels.items(), demangle):
        if not cnt:
            continue
        short = re.sub(r"\(.*", "", dm)
        print(f"{short}")
        print("    " + "  ".join(f"{k} x{v}" for k, v in sorted(cnt.items())))


if __name__ == "__main__":
    sys.exit(main())
