"""N>1 host logic on CPU: world_size-2 gloo.  The SNP-shard plan, the all-gather
of the per-shard frequencies and the ONE all-reduce of the partial phenotypes
 are exercised with the oracle standing in for the per-rank
device work -- no GPU, no product kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from simu_b200.hotpath import _Dist, shard_bounds  # noqa: E402


def test_shard_bounds_balanced_and_contiguous():
    for mc in (0, 1, 7, 8, 100000, 11000000):
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(mc, world)
            assert b[0][0] == 0 and b[-1][1] == mc
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, m, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    rows = oracle.synth_rows(20240901, 0, m, n)
    rng = np.random.default_rng(0)
    causal = np.sort(rng.choice(m, m // 2, replace=False))
    x1, x2 = rng.standard_normal(causal.size), rng.standard_normal(causal.size)
    d = _Dist()
    assert (d.rank, d.world) == (rank, world)
    bounds = shard_bounds(causal.size, world)
    lo, hi = bounds[rank]
    # per-rank work on this rank's slice only
    counts, freq = oracle.find_freq(rows, n, row_index=causal[lo:hi])
    sizes = [b - a for a, b in bounds]
    freq_all = d.all_gather_concat(freq, sizes)
    counts_all = d.all_gather_concat(counts, sizes)
    y1, y2 = oracle.find_pheno(rows, n, freq, x1[lo:hi], x2[lo:hi], row_index=causal[lo:hi])
    y = d.all_reduce_sum(np.ascontiguousarray(np.stack([y1, y2])))
    if rank == 0:
        oc, of = oracle.find_freq(rows, n, row_index=causal)
        o1, o2 = oracle.find_pheno(rows, n, of, x1, x2, row_index=causal)
        np.savez(os.path.join(out_dir, "res.npz"), ok_counts=np.array_equal(counts_all, oc),
                 ok_freq=np.array_equal(freq_all, of),
                 err=max(np.max(np.abs(y[0] - o1)) / np.max(np.abs(o1)), np.max(np.abs(y[1] - o2)) / np.max(np.abs(o2))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_path_matches_single_rank(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), 257, 91, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "res.npz")
    assert bool(r["ok_counts"]) and bool(r["ok_freq"])          # integer work: bit-exact under sharding
    assert float(r["err"]) <= 1e-12                             # sharding only reorders the FP64 sum
