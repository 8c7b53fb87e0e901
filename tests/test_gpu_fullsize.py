"""Full BASELINE sizes on the GPU, checked through size-independent properties
(the CPU oracle would need minutes to hours at these sizes):

* counts: n0+n1+n2+n3 == N for every SNP; freq == (2 n0 + n1) / (2 (n0+n1+n2));
* centring identity: sum_i y_i == 0 up to rounding, because
  sum_i (copies_ij - 2 f_j) [not missing] == (2 n0 + n1) - 2 f_j (n0+n1+n2) == 0;
* FAST vs EXACT mode agree within 1e-5 of max|y| (EXACT is the bit-exact anchor);
* linearity in the effects; a row-subset check against the oracle on a slice
  the oracle finishes in seconds.
"""
import numpy as np
import pytest

from oracle import oracle
from simu_b200 import capi

pytestmark = pytest.mark.gpu
PANEL_SEED = 20240901


def _run(n, m, mc, k2, seed, check_rows=64):
    rng = np.random.default_rng(seed)
    sel = np.sort(rng.choice(m, mc, replace=False)).astype(np.int32) if mc != m else None
    x1 = rng.standard_normal(mc)
    x2 = rng.standard_normal(mc) if k2 else None
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m)
        ctx.panel_synth(PANEL_SEED, 0, m)
        counts, freq = ctx.find_freq(rows=sel, mc=mc)
        assert np.all(counts.sum(axis=1) == n)
        nm = counts[:, 0] + counts[:, 1] + counts[:, 2]
        assert np.array_equal(freq, (2.0 * counts[:, 0] + counts[:, 1]) / (2.0 * nm))
        assert 0.0005 < counts[:, 3].mean() / n < 0.002             # generator: 0.1 % missing
        f1, f2 = ctx.find_pheno(freq, x1, x2, rows=sel, mode=capi.MODE_FAST)
        e1, e2 = ctx.find_pheno(freq, x1, x2, rows=sel, mode=capi.MODE_EXACT)
        g1, _ = ctx.find_pheno(freq, 2.0 * x1, None, rows=sel, mode=capi.MODE_FAST)
        # oracle on the first rows of the causal list (same panel bytes, fetched back from HBM)
        r = check_rows
        sub_rows = np.stack([ctx.panel_get_rows(int(sel[j]) if sel is not None else j, 1)[0] for j in range(r)])
        s1, _ = ctx.find_pheno(freq[:r], x1[:r], None, rows=None if sel is None else sel[:r], mode=capi.MODE_EXACT)
    # the same causal rows as a compact GROUP4 panel (what the CLI stages): identical counts and
    # EXACT phenotypes bit for bit, FAST within the stated tolerance
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_GROUP4)
        if sel is None:
            ctx.panel_synth(PANEL_SEED, 0, mc)
        else:
            ctx.panel_synth_rows(PANEL_SEED, sel.astype(np.int64))
        counts4, freq4 = ctx.find_freq()
        assert np.array_equal(counts4, counts) and np.array_equal(freq4, freq)
        h1, h2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_EXACT)
        assert np.array_equal(h1, e1) and (not k2 or np.array_equal(h2, e2))
        q1, q2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        assert np.max(np.abs(q1 - e1)) <= 1e-5 * np.max(np.abs(e1))
        if k2:
            assert np.max(np.abs(q2 - e2)) <= 1e-5 * np.max(np.abs(e2))
    scale1 = np.max(np.abs(e1))
    assert np.max(np.abs(f1 - e1)) <= 1e-5 * scale1
    if k2:
        assert np.max(np.abs(f2 - e2)) <= 1e-5 * np.max(np.abs(e2))
    assert abs(e1.sum()) <= 1e-9 * np.abs(e1).sum()                  # centring identity
    assert abs(f1.sum()) <= 1e-5 * np.abs(f1).sum()
    assert np.max(np.abs(g1 - 2.0 * f1)) <= 1e-6 * scale1            # linearity (x2 scaling is exact in FP)
    oc, of = oracle.find_freq(sub_rows, n)
    assert np.array_equal(oc, counts[:r]) and np.array_equal(of, freq[:r])
    o1, _ = oracle.find_pheno(sub_rows, n, of, x1[:r])
    assert np.array_equal(s1, o1)


def test_config1_shape():     # 10K x 100K, 1000 causal, 1 trait
    _run(10_000, 100_000, 1_000, False, 1)


def test_config2_shape():     # 100K x 1M, 100K causal, 2 traits (25 GB panel)
    _run(100_000, 1_000_000, 100_000, True, 2)


def test_config3_shape():     # 100K x 1M, 10K causal, 1 trait + liability order
    _run(100_000, 1_000_000, 10_000, False, 3)


def test_config5_shape_biobank():   # 500K samples, 2 traits; 100K-row slice of the 1M-row panel (12.5 GB)
    _run(500_000, 100_000, 20_000, True, 5, check_rows=16)


def test_fully_polygenic_slice():   # config 4 style: every row causal, identity row list
    _run(100_000, 200_000, 200_000, False, 4)
