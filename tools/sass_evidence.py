#!/usr/bin/env python
"""Instruction counts per kernel of the built library + a listing of chosen address ranges (profiles/*_sass_evidence.txt).

usage: python tools/sass_evidence.py [kernel-substring:first-hex:last-hex ...] > profiles/rNN_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "simu_b200", "lib", "libsimu_b200.so")
OPS = ["LDCU", "LDC", "LOP3.LUT", "LDG", "IMAD.HI.U32", "SYNCS", "UTMALDG", "UBLKCP", "LDS", "STS", "SHFL", "ATOMS", "ATOMG", "RED",
       "BAR.SYNC", "BAR.ARV", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "DADD", "DMUL", "POPC", "ACQBULK", "PREEXIT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            kernels[name].append((int(m.group(1), 16), m.group(2).strip()))
    print("SASS evidence (cuobjdump -sass simu_b200/lib/libsimu_b200.so, sm_100a): instruction counts per kernel, then listings of the hot\n"
          "loops.  UTCIMMA = tcgen05.mma kind::i8, STTM / LDTM = tcgen05.st / tcgen05.ld (tensor memory), UTMALDG = cp.async.bulk.tensor (TMA\n"
          "tensor copy), UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit, SYNCS.* = mbarrier operations, LDCU = constant-bank\n"
          "load into a uniform register (the EXACT kernel's effects), ACQBULK / PREEXIT = griddepcontrol.wait / .launch_dependents.\n")
    for k, ins in kernels.items():
        cnt = collections.Counter()
        for _, text in ins:
            op = text.split()[1] if text.startswith("@") else text.split()[0]
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    cnt[o] += 1
                    break
        print(f"{k}  ({len(ins)} instructions)")
        print("    " + ", ".join(f"{o} {c}" for o, c in cnt.items()))
    for spec in sys.argv[1:]:
        sub, lo, hi = spec.split(":")
        lo, hi = int(lo, 16), int(hi, 16)
        for k, ins in kernels.items():
            if sub in k:
                print(f"\n{k}: addresses {lo:04x} .. {hi:04x}\n")
                for a, text in ins:
                    if lo <= a <= hi:
                        print(f"    /*{a:04x}*/  {text}")
                break


if __name__ == "__main__":
    main()
