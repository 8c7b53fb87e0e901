"""BASELINE configs at their FULL sizes on the GPU, compared DIRECTLY with the oracle
(oracle/simu_oracle.c, the restatement of src/source.cc:843-885 and :951-1006).

The oracle's sample loop is split over worker threads (`oracle.find_pheno(threads=...)`): every
sample's sum keeps the reference's ascending-SNP order and operations, so its bits do not change
(tests/test_oracle_golden.py::test_find_pheno_thread_split checks that against the single thread).

* configs 1, 2, 3: every causal row and every sample -- counts and frequencies bit-exact, EXACT
  phenotypes `array_equal`, FAST phenotypes within 1e-5 of max|y| (the tolerance `north_star` states).
* config 5 (500K samples, 100K causal rows, 2 traits) and the per-GPU slice of config 4
  (100K samples, 1,375,000 causal rows): ALL causal rows, on a 2,048-sample column block -- a column
  subset changes nothing about any sample's sum, so the comparison is as strict as above; the columns
  come from the host generator (byte-identical to the device generator: checked here on random rows
  fetched back from HBM).  Counts are compared with the oracle on 512 random rows over all samples.
"""
import os

import numpy as np
import pytest

from oracle import oracle
from simu_b200 import capi

pytestmark = pytest.mark.gpu
PANEL_SEED = 20240901
THREADS = min(os.cpu_count() or 1, 32)


def _tol(a, ref, rel=1e-5):
    assert np.max(np.abs(a - ref)) <= rel * np.max(np.abs(ref)), (np.max(np.abs(a - ref)), np.max(np.abs(ref)))


def _causal(m, mc, seed):
    rng = np.random.default_rng(seed)
    return np.arange(m, dtype=np.int64) if mc == m else np.sort(rng.choice(m, mc, replace=False)).astype(np.int64)


def _effects(mc, k2, seed, freq=None, gcta=False):
    rng = np.random.default_rng(1000 + seed)
    z1, z2 = rng.standard_normal(mc), rng.standard_normal(mc)
    x1, x2 = z1, (0.5 * z1 + np.sqrt(0.75) * z2 if k2 else None)       # --rg 0.5 (source.cc:945-947)
    if gcta:                                                            # --gcta-sigma (source.cc:1257-1262)
        scale = np.sqrt(np.power(np.abs(2.0 * freq * (1.0 - freq)), -1.0))
        x1, x2 = x1 * scale, (x2 * scale if k2 else None)
    return x1, x2


def _full(n, m, mc, k2, seed, gcta=False):
    """every causal row x every sample against the oracle"""
    causal = _causal(m, mc, seed)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        ctx.panel_synth_rows(PANEL_SEED, causal)
        rows = np.empty((mc, ctx.row_bytes), np.uint8)
        step = max(1, (256 << 20) // ctx.row_bytes)
        for lo in range(0, mc, step):
            rows[lo:lo + step] = ctx.panel_get_rows(lo, min(step, mc - lo))
        counts, freq = ctx.find_freq()
        oc, of = oracle.find_freq_threaded(rows, n, THREADS)
        assert np.array_equal(counts, oc) and np.array_equal(freq, of)
        x1, x2 = _effects(mc, k2, seed, freq, gcta)
        o1, o2 = oracle.find_pheno(rows, n, of, x1, x2, threads=THREADS)
        e1, e2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_EXACT)
        assert np.array_equal(e1, o1) and (not k2 or np.array_equal(e2, o2))
        f1, f2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        _tol(f1, o1)
        if k2:
            _tol(f2, o2)
        # the second FAST call on a panel runs with the sparse missing-genotype index instead of the dense MMA plane
        # (csrc/miss_index.cu; panels of >= 2.5e8 genotypes): the same exact fixed-point sums, converted to FP64 in
        # another order -> equal to FP64 rounding
        g1, g2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        assert ctx.missing_index()[0] == (1 if n * mc >= 2.5e8 else 2)
        _tol(g1, f1, 1e-13)
        if k2:
            _tol(g2, f2, 1e-13)


def test_config1_full():      # 10K x 100K, --causal-n 1000, one trait
    _full(10_000, 100_000, 1_000, False, 1)


def test_config2_full():      # 100K x 1M, --causal-n 100000, two traits, --gcta-sigma: 2e10 updates
    _full(100_000, 1_000_000, 100_000, True, 2, gcta=True)


def test_config3_full():      # 100K x 1M, --causal-n 10000, one trait (--cc)
    _full(100_000, 1_000_000, 10_000, False, 3)


def _column_block(n, m, mc, k2, seed, s0, ns=2048, count_rows=512):
    """ALL causal rows on the sample block [s0, s0 + ns) against the oracle"""
    causal = _causal(m, mc, seed)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        if mc == m:
            ctx.panel_synth(PANEL_SEED, 0, m)
        else:
            ctx.panel_synth_rows(PANEL_SEED, causal)
        counts, freq = ctx.find_freq()
        assert np.all(counts.sum(axis=1) == n)
        nm = counts[:, 0] + counts[:, 1] + counts[:, 2]
        assert np.array_equal(freq, (2.0 * counts[:, 0] + counts[:, 1]) / (2.0 * nm))
        # the host generator's columns == the bytes in HBM, and counts == the oracle's, on random rows
        pick = np.sort(np.random.default_rng(seed).choice(mc, count_rows, replace=False))
        dev_rows = np.concatenate([ctx.panel_get_rows(int(j), 1) for j in pick])
        oc, of = oracle.find_freq_threaded(dev_rows, n, THREADS)
        assert np.array_equal(counts[pick], oc) and np.array_equal(freq[pick], of)
        sub = oracle.synth_cols(PANEL_SEED, s0, s0 + ns, row_ids=causal, threads=THREADS)
        assert np.array_equal(sub[pick], dev_rows[:, s0 // 4:(s0 + ns) // 4])
        x1, x2 = _effects(mc, k2, seed)
        o1, o2 = oracle.find_pheno(sub, ns, freq, x1, x2, threads=THREADS)
        e1, e2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_EXACT)
        assert np.array_equal(e1[s0:s0 + ns], o1) and (not k2 or np.array_equal(e2[s0:s0 + ns], o2))
        f1, f2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        _tol(f1, e1)
        _tol(f1[s0:s0 + ns], o1)
        if k2:
            _tol(f2, e2)
            _tol(f2[s0:s0 + ns], o2)
        g1, g2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)      # second use of the panel: sparse missing-genotype index
        assert ctx.missing_index()[0] == 1 and ctx.missing_index()[1] == int(counts[:, 3].sum())
        _tol(g1, f1, 1e-13)                                              # same fixed-point sums, another FP64 conversion order
        if k2:
            _tol(g2, f2, 1e-13)
        assert abs(e1.sum()) <= 1e-9 * np.abs(e1).sum()        # centring identity over ALL samples


def test_config5_all_rows_column_block():      # 500K samples, 100K causal rows of 1M, two traits (12.5 GB)
    _column_block(500_000, 1_000_000, 100_000, True, 5, s0=301_056)


def test_config4_slice_all_rows_column_block():      # one GPU's 1/8 of config 4: 1,375,000 rows, all causal (34 GB)
    _column_block(100_000, 1_375_000, 1_375_000, False, 4, s0=57_344)
