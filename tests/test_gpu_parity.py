"""GPU parity tests proper: the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs.  Bars (north_star):

* allele counts, frequencies, sorted order: bit-exact;
* find_pheno EXACT mode on one GPU: bit-exact (FP64, ordered);
* find_pheno FAST mode: |dy| <= 1e-5 * max|y|   (stated tolerance, FP32 LUT + FP64 flush);
* apply_heritability: <= 1e-12 relative (parallel FP64 sum of squares vs sequential).

Run with `-m gpu` on a B200.
"""
import numpy as np
import pytest

from oracle import oracle
from simu_b200 import capi, plink
from simu_b200.hotpath import HotPath

pytestmark = pytest.mark.gpu

PANEL_SEED = 20240901
FAST_RTOL = 1e-5


def fast_tol(ref, x, mc):
    """FAST mode: 1e-5 of max|y| (north_star's tolerance) plus the absolute floor of the tensor-core kernel's
    fixed-point weights -- each of the mc terms carries at most 2^-41 * max|W| * 3, |W| <= 3 |x| -- which only
    shows when y cancels to (nearly) nothing, e.g. a single sample, whose centred genotypes are all zero."""
    return FAST_RTOL * np.max(np.abs(ref)) + 9.0 * mc * 2.0 ** -41 * (np.max(np.abs(x)) if mc else 0.0)
# every test that does not address rows through a row list runs on both panel layouts
LAYOUTS = pytest.mark.parametrize("layout", [capi.LAYOUT_ROWS, capi.LAYOUT_GROUP4, capi.LAYOUT_SAMPLE16],
                                  ids=["rows", "group4", "sample16"])


def _effects(mc, seed, k2=True):
    rng = np.random.default_rng(seed)
    x1 = rng.standard_normal(mc)
    x2 = 0.5 * x1 + np.sqrt(1 - 0.25) * rng.standard_normal(mc) if k2 else None
    return x1, x2


def _panel(n, m):
    return oracle.synth_rows(PANEL_SEED, 0, m, n)


# ------------------------------------------------------------------ golden

@LAYOUTS
def test_golden_rows_through_the_abi(layout):
    """bed_test.c:237-329 vectors: 0x78 -> [0,1,2,3], 0x2d -> [3,2,1,0]."""
    with capi.Context(4) as ctx:
        ctx.panel_alloc(2, layout)
        ctx.panel_put_rows(0, np.array([[0x78], [0x2D]], np.uint8))
        counts, freq = ctx.find_freq()
        assert counts.tolist() == [[1, 1, 1, 1], [1, 1, 1, 1]]
        assert freq.tolist() == [0.5, 0.5]
        for mode in (capi.MODE_EXACT, capi.MODE_FAST):
            y1, y2 = ctx.find_pheno(freq, [1.0, 10.0], [2.0, -1.0], mode=mode)
            assert y1.tolist() == [1.0, -10.0, -1.0, 10.0]
            assert y2.tolist() == [2.0, 1.0, -2.0, -1.0]


@LAYOUTS
def test_synth_generator_matches_host(layout):
    n, m = 1003, 37
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m, layout)
        ctx.panel_synth(PANEL_SEED, 5, m)
        dev = ctx.panel_get_rows(0, m)
    assert np.array_equal(dev, oracle.synth_rows(PANEL_SEED, 5, m, n))


# ------------------------------------------------------------------ counts

@pytest.mark.parametrize("n", [1, 3, 4, 5, 63, 64, 65, 127, 128, 129, 511, 513, 1000, 4097, 10000, 33333])
@LAYOUTS
def test_counts_bit_exact_ragged_n(n, layout):
    m = 41
    rows = _panel(n, m)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m, layout)
        ctx.panel_put_rows(0, rows)
        counts, freq = ctx.find_freq()
    oc, of = oracle.find_freq(rows, n)
    assert np.array_equal(counts, oc)
    assert np.array_equal(freq, of)          # same IEEE expression -> same bits


@LAYOUTS
def test_counts_ignore_padding_garbage_and_handle_all_missing(layout):
    n = 13
    rows = _panel(n, 6)
    B = plink.row_bytes(n)
    rows[:, B - 1] |= 0xFC                  # garbage in the 3 padding pairs of the last byte
    rows[3, :] = 0x55                       # an all-missing SNP -> freq 0 (source.cc:875)
    wide = np.full((6, B + 9), 0xFF, np.uint8)   # host stride > row bytes, garbage beyond
    wide[:, :B] = rows
    with capi.Context(n) as ctx:
        ctx.panel_alloc(6, layout)
        ctx.panel_put_rows(0, wide)
        counts, freq = ctx.find_freq()
    oc, of = oracle.find_freq(rows, n)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of)
    assert counts[3].tolist() == [0, 0, 0, n] and freq[3] == 0.0


def test_counts_gather_by_row_list():
    n, m = 2500, 300
    rows = _panel(n, m)
    sel = np.sort(np.random.default_rng(1).choice(m, 77, replace=False)).astype(np.int32)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m)
        ctx.panel_put_rows(0, rows)
        counts, freq = ctx.find_freq(rows=sel)
    oc, of = oracle.find_freq(rows, n, row_index=sel)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of)


# -------------------------------------------------------------- find_pheno

@pytest.mark.parametrize("n,mc,k2", [(1, 5, False), (7, 1, True), (64, 64, True), (257, 130, False),
                                     (1000, 333, True), (4099, 257, True), (10000, 1000, False)])
@LAYOUTS
def test_pheno_exact_bit_identical(n, mc, k2, layout):
    rows = _panel(n, mc)
    _, freq = oracle.find_freq(rows, n)
    x1, x2 = _effects(mc, n + mc, k2)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, layout)
        ctx.panel_put_rows(0, rows)
        y1, y2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_EXACT)
    o1, o2 = oracle.find_pheno(rows, n, freq, x1, x2)
    assert np.array_equal(y1, o1)
    if k2:
        assert np.array_equal(y2, o2)


@pytest.mark.parametrize("n,mc,k2", [(1, 5, False), (7, 1, True), (64, 64, True), (257, 130, False),
                                     (1000, 333, True), (4099, 257, True), (10000, 1000, False),
                                     (9001, 4100, True), (20000, 3000, False)])
@LAYOUTS
def test_pheno_fast_within_tolerance(n, mc, k2, layout):
    rows = _panel(n, mc)
    _, freq = oracle.find_freq(rows, n)
    x1, x2 = _effects(mc, 3 * n + mc, k2)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, layout)
        ctx.panel_put_rows(0, rows)
        y1, y2 = ctx.find_pheno(freq, x1, x2, mode=capi.MODE_FAST)
    o1, o2 = oracle.find_pheno(rows, n, freq, x1, x2)
    assert np.max(np.abs(y1 - o1)) <= fast_tol(o1, x1, mc)
    if k2:
        assert np.max(np.abs(y2 - o2)) <= fast_tol(o2, x2, mc)


@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
def test_pheno_gather_missing_and_scaled_effects(mode):
    """--gcta-sigma style effects (x / sqrt(2p(1-p))), a row list, heavy missingness."""
    n, m = 3001, 400
    rows = _panel(n, m)
    rng = np.random.default_rng(5)
    miss = rng.random((m, n)) < 0.2
    codes = np.stack([oracle.unpack(rows[j], n) for j in range(m)])
    codes[miss] = 3
    rows = np.stack([oracle.pack(codes[j]) for j in range(m)])
    sel = np.sort(rng.choice(m, 150, replace=False)).astype(np.int32)
    _, freq = oracle.find_freq(rows, n, row_index=sel)
    het = np.abs(2 * freq * (1 - freq))
    x1 = rng.standard_normal(sel.size) * np.sqrt(het ** -1.0)      # source.cc:1257-1258
    x2 = rng.standard_normal(sel.size) * np.sqrt(het ** -1.0)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m)
        ctx.panel_put_rows(0, rows)
        y1, y2 = ctx.find_pheno(freq, x1, x2, rows=sel, mode=mode)
    o1, o2 = oracle.find_pheno(rows, n, freq, x1, x2, row_index=sel)
    if mode == capi.MODE_EXACT:
        assert np.array_equal(y1, o1) and np.array_equal(y2, o2)
    else:
        assert np.max(np.abs(y1 - o1)) <= FAST_RTOL * np.max(np.abs(o1))
        assert np.max(np.abs(y2 - o2)) <= FAST_RTOL * np.max(np.abs(o2))


@LAYOUTS
def test_pheno_empty_causal_set(layout):
    with capi.Context(10) as ctx:
        ctx.panel_alloc(1, layout)
        y1, y2 = ctx.find_pheno(np.zeros(0), np.zeros(0), np.zeros(0))
        assert not y1.any() and not y2.any()
        counts, freq = ctx.find_freq(mc=0)
        assert counts.shape == (0, 4) and freq.size == 0


@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
@LAYOUTS
def test_pheno_linearity_property_config1_shape(mode, layout):
    """Size-independent property at config-1 shape (10K x 1000 causal):
    y(x_a + x_b) == y(x_a) + y(x_b) and y(c x) == c y(x) up to rounding."""
    n, mc = 10000, 1000
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, layout)
        ctx.panel_synth(PANEL_SEED, 0, mc)
        _, freq = ctx.find_freq()
        xa, xb = _effects(mc, 11, True)
        ya, _ = ctx.find_pheno(freq, xa, mode=mode)
        yb, _ = ctx.find_pheno(freq, xb, mode=mode)
        ys, y3 = ctx.find_pheno(freq, xa + xb, 3.0 * xa, mode=mode)
    scale = np.max(np.abs(ys))
    assert np.max(np.abs(ys - (ya + yb))) <= 2e-5 * scale
    assert np.max(np.abs(y3 - 3.0 * ya)) <= 2e-5 * np.max(np.abs(y3))
    # centred genotypes: sum_i y_i == -sum_j 2 f_j x_j m_j ... for SNPs without missing data the
    # column of centred dosages sums to zero, so |mean(y)| stays tiny next to |y|
    assert abs(ya.mean()) < 0.05 * np.abs(ya).max()


# ------------------------------------------------------------ GROUP4 layout

def test_group4_append_at_any_offset_and_read_back():
    """Rows appended in pieces that do not end on group boundaries (three bfiles of 5, 1 and 13
    causal rows, say) interleave into the same panel as one put; get_rows undoes the layout."""
    n, m = 1001, 19
    rows = _panel(n, m)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m, capi.LAYOUT_GROUP4)
        assert ctx.layout == capi.LAYOUT_GROUP4 and ctx.group_stride == 1024
        for lo, hi in ((0, 5), (5, 6), (6, 19)):
            ctx.panel_put_rows(lo, rows[lo:hi])
        assert np.array_equal(ctx.panel_get_rows(0, m), rows)
        assert np.array_equal(ctx.panel_get_rows(3, 7), rows[3:10])
        counts, freq = ctx.find_freq()
        # overwrite two rows in the middle of groups: the neighbours keep their bits
        ctx.panel_put_rows(6, rows[0:2])
        back = ctx.panel_get_rows(0, m)
    oc, of = oracle.find_freq(rows, n)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of)
    expect = rows.copy(); expect[6:8] = rows[0:2]
    assert np.array_equal(back, expect)


def test_group4_byte_is_the_table_index():
    """Layout contract: byte i of group g = fields of sample i for rows 4g..4g+3, row 4g lowest."""
    import torch
    n, m = 300, 8
    rows = _panel(n, m)
    codes = np.stack([(rows[j, np.arange(n) // 4] >> (2 * (np.arange(n) % 4))) & 3 for j in range(m)])
    host_image = plink.interleave_group4(rows, n)            # the host restatement of the layout
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m, capi.LAYOUT_GROUP4)
        ctx.panel_put_rows(0, rows)
        ctx.sync()
        gs = ctx.group_stride

        class _Mem:
            __cuda_array_interface__ = {"shape": (2, gs), "typestr": "|u1", "data": (ctx.panel_ptr, False), "version": 2,
                                        "strides": None}
        raw = torch.as_tensor(_Mem(), device="cuda").cpu().numpy()
    for g in range(2):
        want = codes[4 * g] | (codes[4 * g + 1] << 2) | (codes[4 * g + 2] << 4) | (codes[4 * g + 3] << 6)
        assert np.array_equal(raw[g, :n], want.astype(np.uint8))
    assert np.array_equal(raw, host_image)


def test_sample16_append_at_any_offset_image_and_row_lists():
    """SAMPLE16 (the CLI's layout): rows appended in pieces that do not end on 16-row group boundaries give the
    same panel as one put; the device image equals the host restatement of the layout (dosage codes, word per
    sample, zero padding); get_rows undoes layout and coding; row lists are refused."""
    import torch
    n, m = 1001, 37
    rows = _panel(n, m)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m, capi.LAYOUT_SAMPLE16)
        assert ctx.layout == capi.LAYOUT_SAMPLE16 and ctx.group_stride == 4 * 1024
        for lo, hi in ((0, 5), (5, 6), (6, 19), (19, 37)):
            ctx.panel_put_rows(lo, rows[lo:hi])
        assert np.array_equal(ctx.panel_get_rows(0, m), rows)
        assert np.array_equal(ctx.panel_get_rows(13, 9), rows[13:22])
        ctx.sync()
        gw = ctx.group_stride // 4

        class _Mem:
            __cuda_array_interface__ = {"shape": (3, gw), "typestr": "<u4", "data": (ctx.panel_ptr, False), "version": 2,
                                        "strides": None}
        raw = torch.as_tensor(_Mem(), device="cuda").cpu().numpy().view(np.uint32)
        assert np.array_equal(raw, plink.interleave_sample16(rows, n))
        counts, freq = ctx.find_freq()
        ctx.panel_put_rows(17, rows[0:2])                     # overwrite two rows inside a group: the neighbours keep their bits
        back = ctx.panel_get_rows(0, m)
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.find_freq(rows=np.array([1, 2], np.int32))
        assert ei.value.status == capi.EINVAL and "SAMPLE16" in str(ei.value)
    oc, of = oracle.find_freq(rows, n)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of)
    expect = rows.copy(); expect[17:19] = rows[0:2]
    assert np.array_equal(back, expect)


def test_group4_rejects_row_lists():
    with capi.Context(100) as ctx:
        ctx.panel_alloc(8, capi.LAYOUT_GROUP4)
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.find_freq(rows=np.array([1, 2], np.int32))
        assert ei.value.status == capi.EINVAL and "GROUP4" in str(ei.value)
        with pytest.raises(capi.SimuB200Error):
            ctx.find_pheno(np.zeros(2), np.zeros(2), rows=np.array([1, 2], np.int32))


@LAYOUTS
def test_synth_rows_scattered(layout):
    n = 777
    ids = np.array([3, 4, 5, 100, 2_000_000, 7], np.int64)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(10, layout)
        ctx.panel_synth_rows(PANEL_SEED, ids, dst_row=3)
        dev = ctx.panel_get_rows(3, ids.size)
    want = np.concatenate([oracle.synth_rows(PANEL_SEED, int(r), 1, n) for r in ids])
    assert np.array_equal(dev, want)


@pytest.mark.parametrize("n,mc", [(100_000, 9), (100_000, 700), (40_000, 5000), (5, 4000)])
def test_group4_counts_warp_split_paths(n, mc):
    """few groups x long rows -> 8/4/2 warps share a group; many groups -> one warp each."""
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_GROUP4)
        ctx.panel_synth(PANEL_SEED, 11, mc)
        counts, freq = ctx.find_freq()
        rows = ctx.panel_get_rows(0, mc)
    oc, of = oracle.find_freq(rows, n)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of)


@pytest.mark.parametrize("which", ["tma", "reg"])
@pytest.mark.parametrize("n,mc", [(100_000, 9), (100_000, 700), (40_000, 5000), (5, 4000), (1, 1), (2049, 33), (8193, 17), (33_333, 257),
                                  (600, 40_000)])
def test_sample16_counts_both_kernels(which, n, mc, monkeypatch):
    """counts_s16_tma_kernel (shared-memory ring, contiguous chunk ranges: whole groups per warp, groups that straddle
    two or more warps' ranges, more warps than chunks) and counts_s16_kernel (register-fed, 1/2/4/8 warps per group)
    -- the library picks by shape, SIMU_B200_COUNTS forces -- both bit-exact, twice in a row (the straddling groups'
    accumulators must be left at zero)."""
    monkeypatch.setenv("SIMU_B200_COUNTS", which)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(mc, capi.LAYOUT_SAMPLE16)
        ctx.panel_synth(PANEL_SEED, 13, mc)
        counts, freq = ctx.find_freq()
        counts2, freq2 = ctx.find_freq()
        part, _ = ctx.find_freq(mc=max(1, mc // 3))                 # a prefix of the panel
        rows = ctx.panel_get_rows(0, mc)
    oc, of = oracle.find_freq_threaded(rows, n, 8)
    assert np.array_equal(counts, oc) and np.array_equal(freq, of)
    assert np.array_equal(counts2, oc) and np.array_equal(freq2, of)
    assert np.array_equal(part, oc[:max(1, mc // 3)])


# ---------------------------------------------------------------- kernel (c)

def test_apply_heritability_parity():
    rng = np.random.default_rng(3)
    n = 12345
    y = rng.standard_normal(n) * 7.0
    noise = rng.standard_normal(n)
    eff = rng.standard_normal(50)
    with capi.Context(n) as ctx:
        for hsq in (0.5, 0.7, 0.0, 1.0):
            gy, ge, gen, env = ctx.apply_heritability(hsq, y, noise, eff)
            oy, oe, ogen, oenv = oracle.apply_heritability(float(np.float32(hsq)), y, noise, eff)
            assert env == oenv
            assert abs(gen - ogen) <= 1e-12 * max(abs(ogen), 1.0)
            assert np.max(np.abs(gy - oy)) <= 1e-12 * np.max(np.abs(oy))
            assert np.max(np.abs(ge - oe)) <= 1e-12 * max(np.max(np.abs(oe)), 1.0)
        # all-zero phenotype: pheno_var <= FLT_EPSILON -> gen 0, env 1 (source.cc:1019-1021)
        gy, _, gen, env = ctx.apply_heritability(0.5, np.zeros(n), noise, None)
        assert (gen, env) == (0.0, 1.0) and np.array_equal(gy, noise)


@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 12345, 151_553, 1_000_003])
def test_apply_heritability_device_forms(n):
    """simu_b200_apply_heritability_dev (one cooperative launch: per-block sums, grid barrier, scaling) against the
    oracle and against simu_b200_sumsq_dev + simu_b200_heritability_dev, twice in a row (the barrier resets itself)."""
    import torch
    rng = np.random.default_rng(n)
    y = rng.standard_normal(n) * 3.0
    noise = rng.standard_normal(n)
    hsq = float(np.float32(0.37))
    oy, _, ogen, oenv = oracle.apply_heritability(hsq, y, noise, None)
    with capi.Context(n) as ctx:
        d_noise = torch.from_numpy(noise).cuda()
        d_coef = torch.zeros(4, dtype=torch.float64, device="cuda")
        for _ in range(2):
            d_y = torch.from_numpy(y).cuda()
            torch.cuda.synchronize()
            ctx.apply_heritability_dev(hsq, d_y.data_ptr(), n, d_noise.data_ptr(), d_coef.data_ptr())
            ctx.sync()
            gy, coef = d_y.cpu().numpy(), d_coef.cpu().numpy()
            assert coef[1] == oenv and abs(coef[0] - ogen) <= 1e-12 * abs(ogen)
            assert np.max(np.abs(gy - oy)) <= 1e-12 * np.max(np.abs(oy))
        d_y2 = torch.from_numpy(y).cuda()
        torch.cuda.synchronize()
        ctx.sumsq_dev(d_y2.data_ptr(), n, d_coef[2:].data_ptr())
        ctx.heritability_dev(hsq, d_y2.data_ptr(), n, d_noise.data_ptr(), d_coef[2:].data_ptr(), 0)
        ctx.sync()
        assert np.max(np.abs(d_y2.cpu().numpy() - gy)) <= 1e-12 * np.max(np.abs(oy))


def test_scale_effects_parity():
    """het^S / sqrt(het) effect scaling (source.cc:1245-1262, :1272-1286, :1089-1094) on the device:
    sqrt and division are correctly rounded -> ops 1, 2 bit-exact; pow -> 1e-14."""
    rng = np.random.default_rng(12)
    mc = 20_011
    freq = rng.uniform(0.001, 0.999, mc)
    x1, x2 = rng.standard_normal(mc), rng.standard_normal(mc)
    comp = rng.integers(0, 3, mc).astype(np.int32)
    sp1, sp2 = np.array([-1.0, -0.25, 0.5]), np.array([-1.0, 0.0, 1.0])
    with capi.Context(16) as ctx:
        for op in (1, 2):
            g1, g2 = ctx.scale_effects(op, freq, x1, x2)
            o1, o2, bad = oracle.scale_effects(op, freq, x1, x2)
            assert bad == -1 and np.array_equal(g1, o1) and np.array_equal(g2, o2)
        g1, g2 = ctx.scale_effects(0, freq, x1, x2, comp, sp1, sp2)
        o1, o2, bad = oracle.scale_effects(0, freq, x1, x2, comp, sp1, sp2)
        assert np.max(np.abs(g1 - o1) / np.abs(o1)) <= 1e-14 and np.max(np.abs(g2 - o2) / np.abs(o2)) <= 1e-14
        g1, _ = ctx.scale_effects(0, freq, x1, None, None, [-1.0])              # --gcta-sigma, one trait
        o1, _, _ = oracle.scale_effects(0, freq, x1, None, None, [-1.0])
        assert np.max(np.abs(g1 - o1) / np.abs(o1)) <= 1e-14
        # a monomorphic causal variant: the reference throws (source.cc:1251-1255)
        f0 = freq.copy(); f0[[777, 5000]] = [0.0, 1.0]
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.scale_effects(0, f0, x1, x2, comp, sp1, sp2)
        assert ei.value.status == capi.ERROR and ctx.last_bad_index == 777
        assert oracle.scale_effects(0, f0, x1, x2, comp, sp1, sp2)[2] == 777
        assert ctx.scale_effects(1, freq[:0], x1[:0])[0].size == 0


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1000, 1024, 1025, 100000])
def test_liability_order_bit_exact(n):
    rng = np.random.default_rng(n)
    y = rng.standard_normal(n)
    if n > 10:
        y[rng.integers(0, n, n // 5)] = 0.25          # ties
        y[rng.integers(0, n, 3)] = -0.0
        y[0], y[1] = np.inf, -np.inf
    with capi.Context(max(n, 4)) as ctx:
        idx = ctx.liability_order(y)
    assert sorted(idx.tolist()) == list(range(n))
    assert np.array_equal(y[idx], np.sort(y))
    # ties resolved by sample index (stable), like the oracle's comparator -- except that
    # -0.0 sorts before +0.0 here; the reference's std::sort leaves all ties unspecified
    if not np.any((y == 0) & np.signbit(y)):
        assert np.array_equal(idx, oracle.liability_order(y))


@pytest.mark.parametrize("n", [1, 2, 31, 33, 1000, 1025, 4097, 16385, 100000, 500001, 700001])
def test_liability_split_matches_the_sorted_order(n):
    """The radix-select split (order statistic) gives the same two pools as ranks [0, k) / [k, n) of
    the stable argsort, for every k incl. the ends, with ties, signed zeros and infinities."""
    rng = np.random.default_rng(n + 1)
    y = rng.standard_normal(n)
    if n > 10:
        y[rng.integers(0, n, n // 4)] = 0.25                       # a big tie class
        y[rng.integers(0, n, n // 50 + 1)] = -1.5
        y[rng.integers(0, n, 3)] = -0.0
        y[rng.integers(0, n, 3)] = 0.0
        y[0], y[1] = np.inf, -np.inf
    with capi.Context(max(n, 4)) as ctx:
        idx = ctx.liability_order(y)
        ks = sorted({0, 1, n // 10, n // 2, max(0, n - 1), n, int(np.searchsorted(np.sort(y), 0.25)) + 7})
        for k in ks:
            if k > n:
                continue
            flags = ctx.liability_split(y, k)
            want = np.ones(n, np.uint8)
            want[idx[:k]] = 0
            assert np.array_equal(flags, want), (n, k)


def test_liability_threshold_default_pools_use_the_split():
    """Default --ncas/--ncon (every sample labelled): the mirror's labels from the split equal the
    oracle's labels from the full sort; the shuffles are still drawn (RNG state)."""
    n, k = 20000, 0.1
    y = np.random.default_rng(4).standard_normal(n)
    calls = []

    def shuffle(lst):
        calls.append(len(lst))
        a = np.array(lst); np.random.default_rng(len(lst)).shuffle(a)
        return a.tolist()

    max_cas, max_con = oracle.liability_split(k, n)
    with capi.Context(n) as ctx:
        lab = HotPath(ctx).apply_liability_threshold(k, max_cas, max_con, shuffle, y)
    assert calls == [max_con, max_cas]
    con = np.arange(max_con); np.random.default_rng(max_con).shuffle(con)
    cas = np.arange(max_con, n); np.random.default_rng(max_cas).shuffle(cas)
    olab = oracle.liability_labels(k, max_cas, max_con, oracle.liability_order(y), con, cas)
    assert np.array_equal(lab, olab) and np.count_nonzero(lab == 2) == max_cas


def test_liability_threshold_labels_agree_with_oracle():
    n = 5000
    rng = np.random.default_rng(9)
    y = rng.standard_normal(n)
    k = 0.1
    perm_rng = np.random.default_rng(1)

    def shuffle(lst):
        a = np.array(lst)
        perm_rng.shuffle(a)
        return a.tolist()

    with capi.Context(n) as ctx:
        hp = HotPath(ctx)
        lab = hp.apply_liability_threshold(k, 300, 700, shuffle, y)
    perm_rng = np.random.default_rng(1)
    max_cas, max_con = oracle.liability_split(k, n)
    con = np.arange(max_con); perm_rng.shuffle(con)
    cas = np.arange(max_con, n); perm_rng.shuffle(cas)
    olab = oracle.liability_labels(k, 300, 700, oracle.liability_order(y), con, cas)
    assert np.array_equal(lab, olab)
    assert (np.count_nonzero(lab == 2), np.count_nonzero(lab == 1)) == (300, 700)


def test_peer_allreduce_single_rank_is_the_identity():
    """World size 1 of the NVLink peer-memory all-reduce (2+ ranks: tools/peer_allreduce_check.py under torchrun)."""
    import torch
    n = 100_003
    with capi.Context(n) as ctx:
        h = ctx.peer_init(0, 1, 2 * n)
        assert len(h) == 64
        ctx.peer_connect([h])
        y = torch.randn(2 * n, dtype=torch.float64, device="cuda")
        want = y.clone()
        for _ in range(5):
            ctx.peer_allreduce_dev(y.data_ptr(), 2 * n)
        ctx.sync()
        torch.cuda.synchronize()
        assert torch.equal(y, want)
        with pytest.raises(capi.SimuB200Error):
            ctx.peer_allreduce_dev(y.data_ptr(), n)            # count differs from peer_init


# ------------------------------------------------------- staging from .bed

@LAYOUTS
def test_load_bed_causal_rows_only(tmp_path, layout):
    n, m = 777, 120
    rows = _panel(n, m)
    prefix = str(tmp_path / "syn")
    plink.write_fileset(prefix, rows, n)
    sel = np.sort(np.random.default_rng(2).choice(m, 40, replace=False))
    with capi.Context(n) as ctx:
        ctx.panel_alloc(40, layout)
        ctx.panel_load_bed(prefix + ".bed", m, sel)
        assert np.array_equal(ctx.panel_get_rows(0, 40), rows[sel])
        counts, _ = ctx.find_freq()
    assert np.array_equal(counts, oracle.find_freq(rows, n, row_index=sel)[0])


def test_load_bed_errors(tmp_path):
    n, m = 50, 10
    rows = _panel(n, m)
    prefix = str(tmp_path / "syn")
    plink.write_fileset(prefix, rows, n)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(m)
        # sample-major header -> "snps must be rows" (source.cc:196-200)
        bad = str(tmp_path / "bad.bed")
        open(bad, "wb").write(bytes([0x6C, 0x1B, 0x00]) + rows[:, :plink.row_bytes(n)].tobytes())
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.panel_load_bed(bad, m)
        assert ei.value.status == capi.EFORMAT and "snps must be rows" in str(ei.value)
        # truncated file -> read error (bed.c:349-352 -> source.cc:226-229)
        short = str(tmp_path / "short.bed")
        open(short, "wb").write(open(prefix + ".bed", "rb").read()[:-7])
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.panel_load_bed(short, m)
        assert ei.value.status == capi.ERROR and "Error reading from" in str(ei.value)
        # .bim announces fewer loci than requested row -> END (source.cc:855-858)
        with pytest.raises(capi.SimuB200Error) as ei:
            ctx.panel_load_bed(prefix + ".bed", m, np.array([3, m]))
        assert ei.value.status == capi.END and "expect to find" in str(ei.value)
        with pytest.raises(capi.SimuB200Error):
            ctx.panel_load_bed(str(tmp_path / "nope.bed"), m)


def test_hotpath_mirror_bfile_chr(tmp_path):
    """--bfile-chr: 3 'chromosomes' concatenated in label order (source.cc:434-443, :277-307);
    the mirror's find_freq / find_pheno against the oracle over the concatenated rows."""
    n = 501
    sizes = [30, 17, 45]
    rows = _panel(n, sum(sizes))
    off = 0
    for lab, sz in zip(("1", "2", "3"), sizes):
        plink.write_fileset(str(tmp_path / f"chr{lab}"), rows[off:off + sz], n, chromosome=int(lab), first_index=off)
        off += sz
    pf = plink.PlinkFiles.from_bfile_chr(str(tmp_path / "chr@"), labels=["1", "2", "3"])
    assert (pf.num_variants, pf.num_samples) == (sum(sizes), n)
    assert pf.get_locus(31).name == "rs31" and pf.get_locus(31).chromosome == 2
    rng = np.random.default_rng(4)
    comp = np.where(rng.random(pf.num_variants) < 0.5, 0, -1)
    causal = np.flatnonzero(comp != -1)
    opts = {"num_variants": pf.num_variants, "num_samples": n, "num_traits": 2}
    x1 = np.zeros(pf.num_variants); x2 = np.zeros(pf.num_variants)
    x1[causal], x2[causal] = _effects(causal.size, 8)
    with capi.Context(n) as ctx:
        hp = HotPath(ctx, pf)
        freq_vec, counts = hp.find_freq(opts, comp, return_counts=True)
        y1, y2 = hp.find_pheno(opts, comp, freq_vec, x1, x2)
        with pytest.raises(RuntimeError, match="expect to find"):
            hp.find_freq(opts, comp[:-1])
    oc, of = oracle.find_freq(rows, n, row_index=causal)
    assert np.array_equal(counts[causal], oc) and np.array_equal(freq_vec[causal], of)
    assert np.all(np.isnan(freq_vec[comp == -1]))                       # source.cc:850
    o1, o2 = oracle.find_pheno(rows, n, of, x1[causal], x2[causal], row_index=causal)
    assert np.array_equal(y1, o1) and np.array_equal(y2, o2)
