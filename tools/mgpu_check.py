#!/usr/bin/env python
"""N-GPU check of the SNP-sharded path through the host mirror (simu_b200/hotpath.py):
    python -m torch.distributed.run --nproc-per-node N tools/mgpu_check.py
Each rank stages ONLY its slice of the causal rows from the .bed files, computes its allele
counts and a partial phenotype; frequencies are all-gathered, phenotypes all-reduced (NCCL).
Rank 0 compares with the CPU oracle on the whole causal set: counts/freq bit-exact, phenotypes
<= 1e-12 (EXACT mode: sharding only reorders the FP64 sum) and <= 1e-5 (FAST mode)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simu_b200 import capi, plink  # noqa: E402
from simu_b200.hotpath import HotPath  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import oracle
    n, sizes = 5003, [700, 450, 850]
    m = sum(sizes)
    rows = oracle.synth_rows(20240901, 0, m, n)
    d = os.path.join(tempfile.gettempdir(), "simu_mgpu_check")
    if rank == 0:
        os.makedirs(d, exist_ok=True)
        off = 0
        for lab, sz in zip("123", sizes):
            plink.write_fileset(os.path.join(d, f"chr{lab}"), rows[off:off + sz], n, chromosome=int(lab), first_index=off)
            off += sz
    dist.barrier()
    pf = plink.PlinkFiles.from_bfile_chr(os.path.join(d, "chr@"), labels=["1", "2", "3"])
    rng = np.random.default_rng(0)
    comp = np.where(rng.random(m) < 0.4, 0, -1)
    causal = np.flatnonzero(comp != -1)
    x1 = np.zeros(m); x2 = np.zeros(m)
    x1[causal] = rng.standard_normal(causal.size); x2[causal] = rng.standard_normal(causal.size)
    opts = {"num_variants": m, "num_samples": n, "num_traits": 2}
    ok = True
    with capi.Context(n, device=local) as ctx:
        hp = HotPath(ctx, pf)
        freq_vec, counts = hp.find_freq(opts, comp, return_counts=True)
        staged = ctx.panel_rows
        res = {}
        for mode, name in ((capi.MODE_EXACT, "exact"), (capi.MODE_FAST, "fast")):
            res[name] = hp.find_pheno(opts, comp, freq_vec, x1, x2, mode=mode)
    if rank == 0:
        oc, of = oracle.find_freq(rows, n, row_index=causal)
        o1, o2 = oracle.find_pheno(rows, n, of, x1[causal], x2[causal], row_index=causal)
        ok &= np.array_equal(counts[causal], oc) and np.array_equal(freq_vec[causal], of)
        err = {k: max(np.max(np.abs(v[0] - o1)) / np.max(np.abs(o1)), np.max(np.abs(v[1] - o2)) / np.max(np.abs(o2)))
               for k, v in res.items()}
        ok &= err["exact"] <= 1e-12 and err["fast"] <= 1e-5
        print(f"mgpu_check world={world}: causal={causal.size}, rows staged on rank 0={staged}, counts/freq bit-exact="
              f"{np.array_equal(counts[causal], oc)}, rel err exact={err['exact']:.2e} fast={err['fast']:.2e} -> "
              f"{'OK' if ok else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
