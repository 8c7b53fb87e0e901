#!/usr/bin/env python
"""Generate tests/golden/ fixtures with the REFERENCE's own compiled libplinkio
(oracle/_ref/libsimu_ref.so, built from /root/reference/libplinkio by oracle/Makefile).

Run in the container that has /root/reference:   python tests/golden/make_golden.py

Writes
  syn_small.bed/.bim/.fam   a 157-sample x 64-SNP synthetic fileset (panel seed 20240901,
                            with an all-missing SNP, a monomorphic SNP and garbage padding bits)
  syn_small_ref.npz         what the reference library + the source.cc loops produce for it:
     codes    [64,157] uint8   pio_next_row output (libplinkio/src/bed.c:321-358)
     comp     [64]     int32   component_per_variant (-1 = non-causal)
     counts   [64,4]   int32   genocount of find_freq (src/source.cc:869-872), causal rows only
     freq     [64]     float64 NaN for non-causal (src/source.cc:850, :875)
     x1, x2   [64]     float64 effect sizes used
     pheno1/2 [157]    float64 find_pheno (src/source.cc:986-996) driven through pio_skip_row / pio_next_row
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from simu_b200 import plink  # noqa: E402


def main():
    R = oracle.ref()
    assert R is not None, "oracle/_ref/libsimu_ref.so missing: run `make -C oracle` where /root/reference exists"
    n, m = 157, 64
    rows = oracle.synth_rows(20240901, 0, m, n)
    B = plink.row_bytes(n)
    rows[5, :] = 0x55                          # all-missing SNP -> freq 0 (source.cc:875)
    rows[9, :] = 0xFF                          # monomorphic hom-A2 SNP -> freq 0
    rows[:, B - 1] |= 0xFC                     # garbage in the 3 padding pairs of the last byte (N % 4 == 1)
    prefix = os.path.join(HERE, "syn_small")
    plink.write_fileset(prefix, rows, n)
    codes = np.zeros((m, n), np.uint8)
    assert R.ref_read_all(prefix.encode(), codes.ctypes.data_as(C.c_void_p)) == m
    rng = np.random.default_rng(42)
    comp = np.where(rng.random(m) < 0.5, 0, -1).astype(np.int32)
    comp[[5, 9, 0, m - 1]] = 0
    x1, x2 = rng.standard_normal(m), rng.standard_normal(m)
    counts = np.zeros((m, 4), np.int32)
    freq = np.zeros(m)
    p1, p2 = np.zeros(n), np.zeros(n)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert R.ref_freq_pheno(prefix.encode(), m, vp(comp), vp(counts), vp(freq), vp(x1), vp(x2), vp(p1), vp(p2), 3) == 0
    np.savez(os.path.join(HERE, "syn_small_ref.npz"), codes=codes, comp=comp, counts=counts, freq=freq, x1=x1, x2=x2,
             pheno1=p1, pheno2=p2)
    print("wrote", prefix + ".{bed,bim,fam}", "and syn_small_ref.npz")


if __name__ == "__main__":
    main()
