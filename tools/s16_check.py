#!/usr/bin/env python
"""SAMPLE16 panel bring-up check on one GPU: staging round trip, counts, EXACT and FAST (tensor-core)
phenotypes against the oracle on awkward shapes.  Usage: python tools/s16_check.py [quick]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from simu_b200 import capi

shapes = [(1, 1), (5, 3), (127, 16), (128, 17), (513, 64), (1000, 65), (4097, 130), (10_000, 1000), (33_333, 257)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    shapes = [(100_000, 20_000)]
bad = 0
for n, mc in shapes:
    rows = oracle.synth_rows(20240901, 7, mc, n)
    rng = np.random.default_rng(n + mc)
    x1, x2 = rng.standard_normal(mc), rng.standard_normal(mc)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        # two pieces at an odd split: boundary groups are read-modify-written
        cut = mc // 3
        if cut:
            ctx.panel_put_rows(0, rows[:cut])
        ctx.panel_put_rows(cut, rows[cut:])
        back = ctx.panel_get_rows(0, mc)
        ok_rt = np.array_equal(back, rows)
        counts, freq = ctx.find_freq()
        oc, of = oracle.find_freq(rows, n)
        ok_c = np.array_equal(counts, oc) and np.array_equal(freq, of)
        o1, o2 = oracle.find_pheno(rows, n, of, x1, x2, threads=8)
        e1, e2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_EXACT)
        ok_e = np.array_equal(e1, o1) and np.array_equal(e2, o2)
        t0 = time.time()
        f1, f2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        dt = time.time() - t0
        s1 = max(np.max(np.abs(o1)), 1e-300); s2 = max(np.max(np.abs(o2)), 1e-300)
        err = max(np.max(np.abs(f1 - o1)) / s1, np.max(np.abs(f2 - o2)) / s2)
        g1, _ = ctx.find_pheno(freq, x1, None, mode=capi.MODE_FAST)
        err1 = np.max(np.abs(g1 - o1)) / s1
        ok_f = err < 1e-9 and err1 < 1e-9
        print(f"N={n} Mc={mc}: roundtrip {ok_rt} counts {ok_c} exact {ok_e} fast K=2 err {err:.2e} K=1 err {err1:.2e} "
              f"({dt*1e3:.1f} ms host call, kernel {ctx.last_kernel_ms(1):.3f} ms)", flush=True)
        if not (ok_rt and ok_c and ok_e and ok_f):
            bad += 1
            if not ok_c:
                w = np.argwhere(counts != oc)
                print("   first count mismatches", w[:5].tolist(), counts[w[0][0]], oc[w[0][0]])
            if not ok_f:
                i = int(np.argmax(np.abs(f1 - o1)))
                print("   worst sample", i, f1[i], o1[i])
print("FAILED" if bad else "ALL OK")
sys.exit(1 if bad else 0)
