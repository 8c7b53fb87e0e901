"""Shape stress for the FAST (LUT) kernel's work split and pipeline: FAST vs EXACT (the
bit-exact anchor, itself checked against the oracle elsewhere) over awkward shapes --
fewer steps than CTAs, one row block, more sample tiles than SMs, row lists with
repeats and descending order, rows near the end of the panel."""
import numpy as np
import pytest

from simu_b200 import capi

pytestmark = pytest.mark.gpu

CASES = [
    # n_samples, panel_rows, mc, traits
    (4096, 70, 1, 1), (4096, 70, 63, 2), (4097, 70, 64, 1), (8191, 200, 65, 2), (8193, 200, 129, 1),
    (12289, 3000, 2999, 2), (40000, 600, 577, 1), (100003, 300, 300, 2), (250001, 130, 128, 1),
    (606208, 80, 64, 1),       # 148 sample tiles = one per SM
    (610305, 70, 70, 2),       # 149 tiles
    (700001, 40, 33, 1),       # more tiles than CTAs
    (1000003, 16, 16, 2),
]


@pytest.mark.parametrize("n,m,mc,k", CASES)
def test_fast_matches_exact(n, m, mc, k):
    rng = np.random.default_rng(n + 7 * m + mc)
    sel = rng.integers(0, m, mc).astype(np.int32)            # repeats allowed, unsorted
    sel[-1] = m - 1
    x1 = rng.standard_normal(mc) * np.exp(rng.standard_normal(mc))     # wide dynamic range of effects
    x2 = rng.standard_normal(mc) if k == 2 else None
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m)
        ctx.panel_synth(99, 0, m)
        counts, freq = ctx.find_freq(rows=sel)
        assert np.all(counts.sum(axis=1) == n)
        e1, e2 = ctx.find_pheno(freq, x1, x2, rows=sel, mode=capi.MODE_EXACT)
        f1, f2 = ctx.find_pheno(freq, x1, x2, rows=sel, mode=capi.MODE_FAST)
        g1, g2 = ctx.find_pheno(freq, x1, x2, rows=sel, mode=capi.MODE_FAST)
    assert np.array_equal(f1, g1)                              # FAST mode is deterministic
    assert np.max(np.abs(f1 - e1)) <= 1e-5 * np.max(np.abs(e1))
    if k == 2:
        assert np.array_equal(f2, g2)
        assert np.max(np.abs(f2 - e2)) <= 1e-5 * np.max(np.abs(e2))


@pytest.mark.parametrize("n,m,mc,k", CASES)
def test_group4_fast_matches_exact(n, m, mc, k):
    """The same shapes on a compact GROUP4 panel (rows 0..mc-1 of the panel, no row list), plus a
    re-staged shorter causal set in the same panel (stale rows behind the last group must not count)."""
    rng = np.random.default_rng(n + 7 * m + mc)
    x1 = rng.standard_normal(mc) * np.exp(rng.standard_normal(mc))
    x2 = rng.standard_normal(mc) if k == 2 else None
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m, capi.LAYOUT_GROUP4)
        ctx.panel_synth(99, 0, m)
        counts, freq = ctx.find_freq(mc=mc)
        assert np.all(counts.sum(axis=1) == n)
        e1, e2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_EXACT)
        f1, f2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
        g1, g2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
    with capi.Context(n) as ctx:                               # row-major reference of the same rows
        ctx.panel_alloc(m)
        ctx.panel_synth(99, 0, m)
        rc, rf = ctx.find_freq(mc=mc)
        r1, r2 = ctx.find_pheno(rf, x1, x2, mode=capi.MODE_EXACT)
    assert np.array_equal(counts, rc) and np.array_equal(freq, rf)
    assert np.array_equal(e1, r1) and (k == 1 or np.array_equal(e2, r2))    # EXACT: same bits in both layouts
    assert np.array_equal(f1, g1)                              # FAST mode is deterministic
    assert np.max(np.abs(f1 - e1)) <= 1e-5 * np.max(np.abs(e1))
    if k == 2:
        assert np.array_equal(f2, g2)
        assert np.max(np.abs(f2 - e2)) <= 1e-5 * np.max(np.abs(e2))


def test_many_contexts_and_reuse():
    """Contexts are independent; scratch grows and is reused across calls of different sizes."""
    a, b = capi.Context(1000), capi.Context(5000)
    try:
        for ctx, n in ((a, 1000), (b, 5000)):
            ctx.panel_alloc(300)
            ctx.panel_synth(5, 0, 300)
        for mc in (300, 7, 150, 300):
            for ctx in (a, b):
                counts, freq = ctx.find_freq(mc=mc)
                x = np.ones(mc)
                e, _ = ctx.find_pheno(freq, x, mode=capi.MODE_EXACT)
                f, _ = ctx.find_pheno(freq, x, mode=capi.MODE_FAST)
                assert np.max(np.abs(e - f)) <= 1e-5 * max(np.max(np.abs(e)), 1e-300)
    finally:
        a.close(); b.close()


@pytest.mark.parametrize("seed", range(14))
def test_sample16_random_shapes_all_paths_against_oracle(seed, monkeypatch):
    """Random SAMPLE16 shapes, missing rates and missingness patterns (none at all, a bad sample, a bad SNP): both counts
    kernels, EXACT (parameter-bank kernel) and FAST with the dense plane and with the sparse missing-genotype index,
    each against the oracle; prefix calls (mc smaller than the panel) included."""
    from oracle import oracle
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 2, 31, 33, 127, 129, 767, 769, 1000, 2049, 4097, 6001]))
    mc = int(rng.choice([1, 15, 16, 17, 63, 65, 127, 129, 200, 513, 1537, 3001]))
    rate = float(rng.choice([0.0, 0.001, 0.003]))
    codes = rng.choice(3, size=(mc, n), p=[0.5, 0.3, 0.2]).astype(np.uint8)       # snp_t codes 0 / 1 / 2 (bed.c:75-85), 3 = missing
    codes[rng.random((mc, n)) < rate] = 3
    if seed % 3 == 1 and n > 8:
        codes[:, n // 2] = 3                                                       # a sample with every genotype missing
    if seed % 3 == 2 and mc > 4:
        codes[mc // 3, :] = 3                                                      # a SNP with every genotype missing
    rows = np.stack([oracle.pack(codes[j]) for j in range(mc)])
    x1 = rng.standard_normal(mc)
    x2 = rng.standard_normal(mc) if seed % 2 else None
    oc, of = oracle.find_freq(rows, n)
    o1, o2 = oracle.find_pheno(rows, n, of, x1, x2)
    scale1 = max(np.max(np.abs(o1)), 1e-300)
    pre = max(1, mc // 2)
    p1, _ = oracle.find_pheno(rows[:pre], n, of[:pre], x1[:pre], None)
    for counts_kernel in ("tma", "reg"):
        monkeypatch.setenv("SIMU_B200_COUNTS", counts_kernel)
        with capi.Context(n) as ctx:
            ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
            ctx.panel_put_rows(0, rows)
            counts, freq = ctx.find_freq()
            assert np.array_equal(counts, oc) and np.array_equal(freq, of)
    monkeypatch.delenv("SIMU_B200_COUNTS")
    for form in ("SIMU_B200_TC_DENSE", "SIMU_B200_TC_SPARSE"):
        monkeypatch.setenv(form, "1")
        with capi.Context(n) as ctx:
            ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
            ctx.panel_put_rows(0, rows)
            e1, e2 = ctx.find_pheno(of, x1, x2, mode=capi.MODE_EXACT)
            assert np.array_equal(e1, o1) and (x2 is None or np.array_equal(e2, o2))
            for _ in range(2):                                                     # the index (if any) is built at the second call
                f1, f2 = ctx.find_pheno(of, x1, x2, mode=capi.MODE_FAST)
                assert np.max(np.abs(f1 - o1)) <= 1e-9 * scale1 + 9.0 * mc * 2.0 ** -41 * np.max(np.abs(x1))
                if x2 is not None:
                    assert np.max(np.abs(f2 - o2)) <= 1e-9 * max(np.max(np.abs(o2)), 1e-300) + 9.0 * mc * 2.0 ** -41 * np.max(np.abs(x2))
            state = ctx.missing_index()[0]
            assert state == 2 if form.endswith("DENSE") else state in (1, 2)       # 2: above the 0.4 % limit (bad sample / SNP on a small panel)
            q1, _ = ctx.find_pheno(of[:pre], x1[:pre], None, mode=capi.MODE_FAST)  # a prefix of the panel
            assert np.max(np.abs(q1 - p1)) <= 1e-9 * max(np.max(np.abs(p1)), 1e-300) + 9.0 * mc * 2.0 ** -41 * np.max(np.abs(x1))
        monkeypatch.delenv(form)
