"""Committed golden fixtures (tests/golden/, produced by the reference's compiled
libplinkio through tests/golden/make_golden.py): the oracle on CPU and the CUDA
path on the GPU must both reproduce them bit for bit."""
import os

import numpy as np
import pytest

from oracle import oracle
from simu_b200 import capi, plink

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    ref = np.load(os.path.join(G, "syn_small_ref.npz"))
    bed = np.fromfile(os.path.join(G, "syn_small.bed"), dtype=np.uint8)
    m, n = ref["codes"].shape
    assert bed[:3].tolist() == [0x6C, 0x1B, 0x01]
    rows = bed[3:].reshape(m, plink.row_bytes(n))
    return ref, rows, n, m


def test_oracle_reproduces_golden():
    ref, rows, n, m = _load()
    for j in range(m):
        assert np.array_equal(oracle.unpack(rows[j], n), ref["codes"][j])
    causal = np.flatnonzero(ref["comp"] != -1)
    counts, freq = oracle.find_freq(rows, n, row_index=causal)
    assert np.array_equal(counts, ref["counts"][causal]) and np.array_equal(freq, ref["freq"][causal])
    assert np.all(np.isnan(ref["freq"][ref["comp"] == -1]))
    assert freq[list(causal).index(5)] == 0.0 and counts[list(causal).index(5)].tolist() == [0, 0, 0, n]
    y1, y2 = oracle.find_pheno(rows, n, freq, ref["x1"][causal], ref["x2"][causal], row_index=causal)
    assert np.array_equal(y1, ref["pheno1"]) and np.array_equal(y2, ref["pheno2"])


def test_host_parsers_on_golden_fileset():
    pf = plink.PlinkFiles([os.path.join(G, "syn_small")])
    assert (pf.num_samples, pf.num_variants) == (157, 64)
    assert pf.get_locus(63).name == "rs63" and pf.get_sample(156).iid == "id2_156"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
def test_cuda_path_reproduces_golden(mode):
    ref, rows, n, m = _load()
    causal = np.flatnonzero(ref["comp"] != -1)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(causal.size)
        ctx.panel_load_bed(os.path.join(G, "syn_small.bed"), m, causal)          # staged straight from the fixture file
        counts, freq = ctx.find_freq()
        assert np.array_equal(counts, ref["counts"][causal]) and np.array_equal(freq, ref["freq"][causal])
        y1, y2 = ctx.find_pheno(freq, ref["x1"][causal], ref["x2"][causal], mode=mode)
    if mode == capi.MODE_EXACT:
        assert np.array_equal(y1, ref["pheno1"]) and np.array_equal(y2, ref["pheno2"])
    else:
        assert np.max(np.abs(y1 - ref["pheno1"])) <= 1e-5 * np.max(np.abs(ref["pheno1"]))
        assert np.max(np.abs(y2 - ref["pheno2"])) <= 1e-5 * np.max(np.abs(ref["pheno2"]))
