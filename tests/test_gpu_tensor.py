"""The tensor-core find_pheno kernel (SAMPLE16 panels, MODE_FAST; csrc/pheno_tc.cu) against the oracle
(src/source.cc:951-1006).  The kernel computes the EXACT integer sum of per-SNP weights quantised to 40-bit
fixed point relative to the largest weight of the trait, so its error is set by that quantisation alone:
the tests state 1e-9 of max|y| where the effect sizes are well scaled (the tolerance north_star asks for is
1e-5) and check the 2^-41 max|W| per-term bound where they are not."""
import numpy as np
import pytest

from oracle import oracle
from simu_b200 import capi

pytestmark = pytest.mark.gpu


def _run(rows, n, x1, x2, freq=None):
    mc = rows.shape[0]
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        ctx.panel_put_rows(0, rows)
        counts, f = ctx.find_freq()
        y = ctx.find_pheno(f if freq is None else freq, x1, x2, mode=capi.MODE_FAST)
    return counts, f, y


# shapes around every granularity of the kernel: 128-sample sub-tiles, 768-sample tiles, 16-row groups,
# 128-SNP stages, the 16384-SNP drain interval, more steps than SMs, fewer steps than SMs
@pytest.mark.parametrize("n,mc,k2", [(1, 1, True), (127, 15, False), (129, 17, True), (767, 127, True), (768, 128, False),
                                     (769, 129, True), (1537, 2049, False), (5003, 16384, True), (777, 16400, False),
                                     (2000, 40000, True), (33333, 257, True), (100003, 300, False)])
def test_tensor_kernel_matches_oracle(n, mc, k2):
    rows = oracle.synth_rows(20240901, 11, mc, n)
    rng = np.random.default_rng(n * 7 + mc)
    x1 = rng.standard_normal(mc)
    x2 = 0.5 * x1 + np.sqrt(0.75) * rng.standard_normal(mc) if k2 else None
    counts, freq, (y1, y2) = _run(rows, n, x1, x2)
    oc, of = oracle.find_freq_threaded(rows, n, 8)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of)
    o1, o2 = oracle.find_pheno(rows, n, of, x1, x2, threads=8)
    # 1e-9 of max|y|, plus the fixed-point floor (each term carries <= 2^-41 max|W| * 3 with |W| <= 3 |x|), which
    # only shows where y cancels to nothing: one sample's centred genotypes are all zero
    assert np.max(np.abs(y1 - o1)) <= 1e-9 * np.max(np.abs(o1)) + 9.0 * mc * 2.0 ** -41 * np.max(np.abs(x1))
    if k2:
        assert np.max(np.abs(y2 - o2)) <= 1e-9 * np.max(np.abs(o2)) + 9.0 * mc * 2.0 ** -41 * np.max(np.abs(x2))


def test_tensor_kernel_heavy_missingness_and_all_missing_rows():
    n, mc = 3001, 400
    rng = np.random.default_rng(5)
    rows = oracle.synth_rows(3, 0, mc, n)
    codes = np.stack([oracle.unpack(rows[j], n) for j in range(mc)])
    codes[rng.random((mc, n)) < 0.3] = 3
    codes[7] = 3                                                   # a row with every genotype missing: freq 0 (source.cc:875)
    codes[100] = 0                                                 # monomorphic
    rows = np.stack([oracle.pack(codes[j]) for j in range(mc)])
    x1, x2 = rng.standard_normal(mc), rng.standard_normal(mc)
    counts, freq, (y1, y2) = _run(rows, n, x1, x2)
    oc, of = oracle.find_freq(rows, n)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of) and freq[7] == 0.0 and counts[7, 3] == n
    o1, o2 = oracle.find_pheno(rows, n, of, x1, x2)
    assert np.max(np.abs(y1 - o1)) <= 1e-9 * np.max(np.abs(o1))
    assert np.max(np.abs(y2 - o2)) <= 1e-9 * np.max(np.abs(o2))


def test_tensor_kernel_quantisation_bound_with_badly_scaled_effects():
    """One effect 1e9 times the others: the small ones are represented to 2^-41 of the LARGE one, which is the
    documented bound (error per term <= 2^-41 max|W| * 3); trait 2 has its own scale and stays tight."""
    n, mc = 2048, 1000
    rows = oracle.synth_rows(9, 0, mc, n)
    rng = np.random.default_rng(1)
    x1 = rng.standard_normal(mc)
    x1[500] = 1e9
    x2 = rng.standard_normal(mc)
    _, freq, (y1, y2) = _run(rows, n, x1, x2)
    o1, o2 = oracle.find_pheno(rows, n, freq, x1, x2)
    wmax = 1e9 * 3.0                                               # |x (1 + 2f)| <= 3 |x|
    assert np.max(np.abs(y1 - o1)) <= mc * 3 * wmax * 2.0 ** -41 + 1e-9 * np.max(np.abs(o1))
    assert np.max(np.abs(y2 - o2)) <= 1e-9 * np.max(np.abs(o2))


def test_tensor_kernel_zero_effects_and_is_deterministic():
    n, mc = 5000, 3000
    rows = oracle.synth_rows(2, 0, mc, n)
    rng = np.random.default_rng(2)
    x1 = rng.standard_normal(mc)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        ctx.panel_put_rows(0, rows)
        _, freq = ctx.find_freq()
        z1, z2 = ctx.find_pheno(freq, np.zeros(mc), np.zeros(mc), mode=capi.MODE_FAST)
        assert not z1.any() and not z2.any()
        a1, a2 = ctx.find_pheno(freq, x1, -x1, mode=capi.MODE_FAST)
        b1, b2 = ctx.find_pheno(freq, x1, -x1, mode=capi.MODE_FAST)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)      # integer sums + fixed slot order: same bits every run
    assert np.array_equal(a2, -a1)                                 # sign symmetry of the fixed-point weights


def test_sparse_missing_index_and_dense_plane_agree():
    """The missing genotypes as a list (csrc/miss_index.cu: one correction pass of exact 64-bit integer sums) and as the
    dense m plane of the MMA are the exact sum of the same quantised weights; only the points where integers become
    FP64 differ, so the results agree to FP64 rounding (1e-13 of max|y| stated; identical bits on this small case).  Also
    checks which form runs: the list below 0.4 % missing, the plane above, and again the list after the panel is
    rewritten with other rows (the index follows the panel)."""
    import os
    os.environ["SIMU_B200_TC_SPARSE"] = "1"                        # small panels default to the dense plane (launch-bound)
    n, mc = 5003, 3000
    rows = oracle.synth_rows(20240901, 21, mc, n)                  # 0.1 % missing
    codes = np.stack([oracle.unpack(rows[j], n) for j in range(mc)])
    codes[:2500, 17] = 3                                           # one bad sample: more entries than one warp of the pass takes
    rows = np.stack([oracle.pack(codes[j]) for j in range(mc)])
    rng = np.random.default_rng(21)
    x1, x2 = rng.standard_normal(mc), rng.standard_normal(mc)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        ctx.panel_put_rows(0, rows)
        _, freq = ctx.find_freq()
        assert ctx.missing_index()[0] == 0
        first1, first2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)          # first use of a panel: dense plane, no index yet
        assert ctx.missing_index()[0] == 0
        s1, s2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)                  # second use: the index is built
        assert np.max(np.abs(first1 - s1)) <= 1e-13 * np.max(np.abs(s1)) and np.max(np.abs(first2 - s2)) <= 1e-13 * np.max(np.abs(s2))
        state, entries = ctx.missing_index()
        assert state == 1 and entries == int((codes == 3).sum())
        t1, _ = ctx.find_pheno(freq, x1, None, mode=capi.MODE_FAST)                   # one trait through the same index
        u1, u2 = ctx.find_pheno(freq[:1000], x1[:1000], x2[:1000], mode=capi.MODE_FAST)   # a prefix of the panel: later rows' entries are skipped
    os.environ["SIMU_B200_TC_DENSE"] = "1"
    try:
        with capi.Context(n) as ctx:
            ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
            ctx.panel_put_rows(0, rows)
            d1, d2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
            assert ctx.missing_index()[0] == 2
            e1, _ = ctx.find_pheno(freq, x1, None, mode=capi.MODE_FAST)
            v1, v2 = ctx.find_pheno(freq[:1000], x1[:1000], x2[:1000], mode=capi.MODE_FAST)
    finally:
        del os.environ["SIMU_B200_TC_DENSE"]
    for a, b in ((s1, d1), (s2, d2), (t1, e1), (u1, v1), (u2, v2)):
        assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b))
    o1, o2 = oracle.find_pheno(rows, n, freq, x1, x2, threads=8)
    assert np.max(np.abs(s1 - o1)) <= 1e-9 * np.max(np.abs(o1)) and np.max(np.abs(s2 - o2)) <= 1e-9 * np.max(np.abs(o2))
    # above the rate limit the plane is used; rewriting the panel rebuilds the index
    dense_rows = rows.copy()
    codes[rng.random((mc, n)) < 0.02] = 3
    dense_rows = np.stack([oracle.pack(codes[j]) for j in range(mc)])
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        ctx.panel_put_rows(0, dense_rows)
        _, f2 = ctx.find_freq()
        h1, h2 = ctx.find_pheno(f2, x1, x2, mode=capi.MODE_FAST)
        h1, h2 = ctx.find_pheno(f2, x1, x2, mode=capi.MODE_FAST)
        assert ctx.missing_index()[0] == 2
        p1, p2 = oracle.find_pheno(dense_rows, n, f2, x1, x2, threads=8)
        assert np.max(np.abs(h1 - p1)) <= 1e-9 * np.max(np.abs(p1)) and np.max(np.abs(h2 - p2)) <= 1e-9 * np.max(np.abs(p2))
        ctx.panel_put_rows(0, rows)                                # back to the sparse panel: the index is rebuilt
        assert ctx.missing_index()[0] == 0
        r1, r2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        r1, r2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        assert ctx.missing_index()[0] == 1
        assert np.array_equal(r1, s1) and np.array_equal(r2, s2)             # the same path twice: the same bits
    del os.environ["SIMU_B200_TC_SPARSE"]
