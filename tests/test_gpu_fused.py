"""simu_b200_find_freq_pheno (counts + effect scaling + FAST phenotypes behind one call, compact panels of
both layouts) against the calls it replaces:
    find_freq -> scale_effects -> find_pheno
counts and frequencies bit-exact, scaled effects within 1e-14, phenotypes within the FAST
tolerance of the bit-exact EXACT mode (which the parity tests pin against the oracle)."""
import numpy as np
import pytest

from simu_b200 import capi

pytestmark = pytest.mark.gpu

CASES = [(4096, 64, False), (4096, 64, True), (5003, 257, True), (10000, 1000, False), (100003, 300, True),
         (20000, 3000, False), (1, 5, False), (7, 1, True), (250001, 130, False), (500000, 80, False), (40000, 5000, True)]


@pytest.fixture(params=[capi.LAYOUT_GROUP4, capi.LAYOUT_SAMPLE16], ids=["group4", "sample16"])
def impl(request):
    return request.param


@pytest.mark.parametrize("n,mc,k2", CASES)
def test_one_call_matches_the_three_calls(n, mc, k2, impl):
    rng = np.random.default_rng(n + mc)
    r1 = rng.standard_normal(mc)
    r2 = rng.standard_normal(mc) if k2 else None
    comp = rng.integers(0, 2, mc).astype(np.int32)
    sp1, sp2 = np.array([-1.0, 0.5]), np.array([-0.25, -1.0])
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, impl)
        ctx.panel_synth(99, 0, mc)
        counts, freq = ctx.find_freq()
        for op in ((-1, 0, 1) if n > 100 else (-1,)):          # tiny N: monomorphic variants -> zero het
            x1, x2 = ctx.scale_effects(op, freq, r1, r2, comp, sp1, sp2) if op >= 0 else (r1, r2)
            e1, e2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_EXACT)
            c, f, g1, g2, y1, y2 = ctx.find_freq_pheno(r1, r2, op, comp, sp1, sp2)
            assert np.array_equal(c, counts) and np.array_equal(f, freq)
            assert np.allclose(g1, x1, rtol=1e-14, atol=0)
            # + the fixed-point floor of the tensor-core kernel (2^-41 max|W| per term), which shows when y cancels to ~0
            assert np.max(np.abs(y1 - e1)) <= 1e-5 * np.max(np.abs(e1)) + 9.0 * mc * 2.0 ** -41 * np.max(np.abs(x1))
            if k2:
                assert np.allclose(g2, x2, rtol=1e-14, atol=0)
                assert np.max(np.abs(y2 - e2)) <= 1e-5 * np.max(np.abs(e2)) + 9.0 * mc * 2.0 ** -41 * np.max(np.abs(x2))


def test_one_call_reports_zero_het_and_needs_a_compact_panel(impl):
    n, mc = 3000, 40
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, impl)
        rows = np.zeros((mc, (n + 3) // 4), np.uint8)           # all hom A1 -> freq 1 -> het 0
        rows[5:] = ctx_rows = np.random.default_rng(0).integers(0, 256, (mc - 5, (n + 3) // 4), dtype=np.uint8)
        ctx.panel_put_rows(0, rows)
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.find_freq_pheno(np.ones(mc), None, 0, None, [-1.0])
        assert ei.value.status == capi.ERROR and ctx.last_bad_index == 0
        c, f, x, _, y, _ = ctx.find_freq_pheno(np.ones(mc), None, -1)        # no scaling: no error
        assert f[0] == 1.0 and c[0].tolist() == [n, 0, 0, 0]
        assert ctx.find_freq_pheno(np.zeros(0), None, -1)[4].shape == (n,)   # empty causal set -> zeros
        ctx.panel_alloc(mc)                                                   # ROWS panel
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.find_freq_pheno(np.ones(mc), None, -1)
        assert ei.value.status == capi.EINVAL
