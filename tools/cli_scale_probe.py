#!/usr/bin/env python
"""Wall time of the drop-in CLI at scale, per stage (SIMU_B200_TIMING=1), on a synthetic bfile that
is generated on the GPU and written to local disk:  python tools/cli_scale_probe.py [N] [M] [causal_n]"""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simu_b200 import capi, plink

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
mc = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
d = tempfile.mkdtemp(prefix="simu_cli_")
prefix = os.path.join(d, "syn")
t0 = time.time()
with capi.Context(n) as ctx, open(prefix + ".bed", "wb") as f:
    f.write(plink.BED_MAGIC)
    chunk = 20_000
    ctx.panel_alloc(chunk)
    for lo in range(0, m, chunk):
        nr = min(chunk, m - lo)
        ctx.panel_synth(20240901, lo, nr)
        f.write(ctx.panel_get_rows(0, nr).tobytes())
plink.write_bim(prefix + ".bim", m)
plink.write_fam(prefix + ".fam", n)
print(f"fileset {n} x {m}: {os.path.getsize(prefix + '.bed') / 1e9:.2f} GB .bed written in {time.time() - t0:.1f} s", flush=True)
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "simu_b200", "bin", "simu_b200")
for name, extra in (("qt K=2 gcta", ["--qt", "--num-traits", "2", "--rg", "0.5", "--gcta-sigma", "--norm-effect", "--hsq", "0.5", "0.7"]),
                    ("cc K=1", ["--cc", "--k", "0.1", "--hsq", "0.5"])):
    for rep in range(4):
        t0 = time.time()
        r = subprocess.run([cli, "--bfile", prefix, *extra, "--causal-n", str(mc), "--seed", "1", "--out",
                            os.path.join(d, "out")], env=dict(os.environ, SIMU_B200_TIMING="1"), capture_output=True, text=True)
        print(f"{name} run {rep}: rc={r.returncode} wall {time.time() - t0:.2f} s  {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}", flush=True)
