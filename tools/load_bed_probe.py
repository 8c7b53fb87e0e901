"""Measure simu_b200_panel_load_bed on a synthetic fileset (page-cache warm): GB/s for all rows
and for a scattered 10 % causal subset, with 1 and 8 reader threads."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simu_b200 import capi, plink

n, m = 100_000, 40_000          # 1 GB .bed
B = (n + 3) // 4
path = "/tmp/probe.bed"
with capi.Context(n) as ctx:
    ctx.panel_alloc(m)
    ctx.panel_synth(20240901, 0, m)
    rows = ctx.panel_get_rows(0, m)
    plink.write_bed(path, rows, n)
    sel = np.sort(np.random.default_rng(0).choice(m, m // 10, replace=False))
    for thr in (1, 8):
        os.environ["SIMU_B200_IO_THREADS"] = str(thr)
        for name, fr in (("all rows", None), ("10% scattered", sel)):
            k = m if fr is None else fr.size
            ctx.panel_load_bed(path, m, fr); ctx.sync()
            t = time.perf_counter(); ctx.panel_load_bed(path, m, fr); ctx.sync(); dt = time.perf_counter() - t
            ok = np.array_equal(ctx.panel_get_rows(0, 16), rows[:16] if fr is None else rows[fr[:16]])
            print(f"threads={thr} {name}: {k * B / dt / 1e9:.2f} GB/s ok={ok}")
os.remove(path)
