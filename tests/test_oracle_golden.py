"""Pin the CPU oracle against every known-answer vector the reference holds for
this path and against the reference's own compiled libplinkio.

Golden vectors: /root/reference/libplinkio/tests/bed_test.c:144-329.
CPU only -- no GPU, no product code.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from simu_b200 import plink  # noqa: E402


# ---- bed_test.c:237-250 test_unpack_snps
def test_unpack_0x78():
    assert oracle.unpack(np.array([0x78], np.uint8), 4).tolist() == [0, 1, 2, 3]


# ---- bed_test.c:297-329 test_bed_skip_row (second row 0x2d -> [3,2,1,0])
def test_unpack_0x2d():
    assert oracle.unpack(np.array([0x2D], np.uint8), 4).tolist() == [3, 2, 1, 0]


def test_pack_roundtrip_and_bits():
    # snp_lookup.h:27  snp_to_bits = {0, 2, 3, 1}
    assert oracle.pack(np.array([0, 1, 2, 3], np.uint8)).tolist() == [0x78]
    assert oracle.pack(np.array([3, 2, 1, 0], np.uint8)).tolist() == [0x2D]
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4, 5, 7, 8, 63, 64, 65, 1001):
        codes = rng.integers(0, 4, n).astype(np.uint8)
        assert np.array_equal(oracle.unpack(oracle.pack(codes), n), codes)


# ---- bed_test.c:144-193 header cases
def test_header_v100_snp_major():
    assert oracle.bed_header(bytes([0x6C, 0x1B, 0x01])) == (100, 1, 3)


def test_header_v100_sample_major():      # bed_test.c:217-235 test_bed_open2
    assert oracle.bed_header(bytes([0x6C, 0x1B, 0x00])) == (100, 0, 3)


def test_header_v099():                   # bed_test.c:160-176
    assert oracle.bed_header(bytes([0, 0, 0])) == (99, 0, 1)


def test_row_bytes():
    # bed_header.c:219-223 (the stale expectation of bed_test.c:252-266 is about the
    # unpacked size and is not in the reference's run list)
    assert [oracle.row_bytes(n) for n in (1, 4, 5, 7, 8, 100000)] == [1, 1, 2, 2, 2, 25000]


# ---- hand-computed find_freq / find_pheno on the golden rows
def test_find_freq_golden_rows():
    rows = np.array([[0x78], [0x2D]], np.uint8)
    counts, freq = oracle.find_freq(rows, 4)
    assert counts.tolist() == [[1, 1, 1, 1], [1, 1, 1, 1]]
    assert freq.tolist() == [0.5, 0.5]          # (2*1+1)/(2*3)


def test_find_pheno_golden_rows():
    rows = np.array([[0x78], [0x2D]], np.uint8)
    _, freq = oracle.find_freq(rows, 4)
    y1, y2 = oracle.find_pheno(rows, 4, freq, [1.0, 10.0], [2.0, -1.0])
    # row 0 codes [0,1,2,3] -> copies [2,1,0,missing]; row 1 codes [3,2,1,0]
    # centred by 2f = 1: row0 -> [1,0,-1,0] * x ; row1 -> [0,-1,0,1] * x
    assert y1.tolist() == [1.0, -10.0, -1.0, 10.0]
    assert y2.tolist() == [2.0, 1.0, -2.0, -1.0]


def test_all_missing_row_has_zero_freq():   # source.cc:875
    rows = np.array([[0x55, 0x55]], np.uint8)
    counts, freq = oracle.find_freq(rows, 7)
    assert counts.tolist() == [[0, 0, 0, 7]] and freq.tolist() == [0.0]


def test_padding_bits_are_ignored():
    # N = 5: the 3 padding pairs of byte 1 hold garbage; source.cc:870 only walks N samples
    clean = np.array([[0x78, 0x02]], np.uint8)
    dirty = np.array([[0x78, 0xFE]], np.uint8)
    a, b = oracle.find_freq(clean, 5), oracle.find_freq(dirty, 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_apply_heritability_formula():
    y = np.array([1.0, -2.0, 3.0, 0.5])
    noise = np.array([0.1, 0.2, -0.3, 0.4])
    eff = np.array([1.0, 2.0])
    hsq = np.float32(0.7)
    out, e, gen, env = oracle.apply_heritability(float(hsq), y, noise, eff)
    var = float(np.sum(y * y) / 4.0)
    assert gen == np.sqrt(float(hsq) / var) and env == np.sqrt(1.0 - float(hsq))
    assert np.array_equal(out, y * gen + noise * env)
    assert np.array_equal(e, eff * gen)
    # zero heritability: gen 0, env 1 (source.cc:1019-1021)
    out, e, gen, env = oracle.apply_heritability(0.0, y, noise, eff)
    assert (gen, env) == (0.0, 1.0) and np.array_equal(out, noise)


def test_liability_split_uses_float_product():   # source.cc:1038
    assert oracle.liability_split(0.1, 100000) == (10000, 90000)
    k = np.float32(0.3)
    n = 10
    assert oracle.liability_split(float(k), n)[0] == int(np.floor(np.float32(k) * np.float32(n)))


def test_liability_labels_default_counts():
    y = np.array([0.3, -1.0, 2.0, 0.1, 5.0, -0.2, 0.0, 1.0, 0.9, -3.0])
    idx = oracle.liability_order(y)
    assert np.array_equal(y[idx], np.sort(y))
    max_cas, max_con = oracle.liability_split(0.2, y.size)
    lab = oracle.liability_labels(0.2, max_cas, max_con, idx, np.arange(max_con), np.arange(max_con, y.size))
    assert sorted(np.flatnonzero(lab == 2).tolist()) == sorted(np.argsort(y)[-2:].tolist())
    assert np.count_nonzero(lab == 1) == 8


# ---- the restated decode agrees with the reference's compiled libplinkio
@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built (no /root/reference and no prebuilt .so)")
@pytest.mark.parametrize("n,m", [(4, 2), (7, 5), (64, 3), (65, 9), (1001, 17)])
def test_oracle_vs_reference_libplinkio(tmp_path, n, m):
    import ctypes as C
    rows = oracle.synth_rows(20240901, 0, m, n)
    prefix = str(tmp_path / "syn")
    plink.write_fileset(prefix, rows, n)
    R = oracle.ref()
    ns, nl, lpr = C.c_int64(), C.c_int64(), C.c_int()
    assert R.ref_dims(prefix.encode(), C.byref(ns), C.byref(nl), C.byref(lpr)) == 0
    assert (ns.value, nl.value, lpr.value) == (n, m, 1)
    out = np.zeros((m, n), np.uint8)
    assert R.ref_read_all(prefix.encode(), out.ctypes.data_as(C.c_void_p)) == m
    mine = np.stack([oracle.unpack(rows[j], n) for j in range(m)])
    assert np.array_equal(out, mine)

    # and the restated find_freq / find_pheno agree bit for bit with the same loops driven
    # through pio_skip_row / pio_next_row on the real library
    rng = np.random.default_rng(n * 131 + m)
    comp = np.where(rng.random(m) < 0.6, 0, -1).astype(np.int32)
    comp[0] = 0
    causal = np.flatnonzero(comp != -1)
    x1 = rng.standard_normal(m)
    x2 = rng.standard_normal(m)
    counts = np.zeros((m, 4), np.int32)
    freq = np.zeros(m)
    p1, p2 = np.zeros(n), np.zeros(n)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert R.ref_freq_pheno(prefix.encode(), m, vp(comp), vp(counts), vp(freq), vp(x1), vp(x2), vp(p1), vp(p2), 3) == 0
    oc, of = oracle.find_freq(rows, n, row_index=causal)
    assert np.array_equal(oc, counts[causal]) and np.array_equal(of, freq[causal])
    assert np.all(np.isnan(freq[comp == -1]))
    y1, y2 = oracle.find_pheno(rows, n, of, x1[causal], x2[causal], row_index=causal)
    assert np.array_equal(y1, p1) and np.array_equal(y2, p2)


@pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built")
def test_reference_parsers_on_their_own_vectors(tmp_path):
    """bim_test.c:50-77 and fam_test.c:87-115 strings through the real parser and ours."""
    import ctypes as C
    prefix = str(tmp_path / "v")
    open(prefix + ".bim", "w").write("1 rs1 0 1234567 A C\n1 rs2 0.23 7654321 - ACCG\n")
    open(prefix + ".fam", "w").write("F1 P1 0 0 1 1\nF1\t P2 0 0 2 2\n")
    plink.write_bed(prefix + ".bed", np.array([[0x08], [0x0D]], np.uint8), 2)
    R = oracle.ref()
    mine = plink.PlinkFile.open(prefix)
    assert (mine.num_samples, mine.num_loci) == (2, 2)
    for i, exp in enumerate([(1, "rs1", 1234567, "A", "C"), (1, "rs2", 7654321, "-", "ACCG")]):
        ch, bp = C.c_int(), C.c_longlong()
        name, a1, a2 = C.create_string_buffer(64), C.create_string_buffer(16), C.create_string_buffer(16)
        assert R.ref_locus(prefix.encode(), i, C.byref(ch), name, 64, C.byref(bp), a1, a2, 16) == 0
        got = (ch.value, name.value.decode(), bp.value, a1.value.decode(), a2.value.decode())
        assert got == exp
        lo = mine.loci[i]
        assert (lo.chromosome, lo.name, lo.bp_position, lo.allele1, lo.allele2) == exp
    for i, exp in enumerate([("F1", "P1"), ("F1", "P2")]):
        fid, iid = C.create_string_buffer(64), C.create_string_buffer(64)
        assert R.ref_sample(prefix.encode(), i, fid, iid, 64) == 0
        assert (fid.value.decode(), iid.value.decode()) == exp
        assert (mine.samples[i].fid, mine.samples[i].iid) == exp


def test_synth_rows_statistics():
    n, m = 20000, 40
    rows = oracle.synth_rows(20240901, 0, m, n)
    counts, freq = oracle.find_freq(rows, n)
    assert counts.sum(axis=1).tolist() == [n] * m
    assert 0.0 < counts[:, 3].mean() / n < 0.003          # ~0.1% missing
    assert freq.min() > 0.0 and freq.max() < 0.55          # p ~ U(0.01, 0.5)
    # deterministic and row-addressable
    again = oracle.synth_rows(20240901, 7, 3, n)
    assert np.array_equal(again, rows[7:10])


def test_scale_effects_restatement():
    """oracle.scale_effects against the formulas of source.cc:1250-1262, :1277-1285, :1092-1093."""
    rng = np.random.default_rng(3)
    f = rng.uniform(0.01, 0.99, 500)
    x1, x2 = rng.standard_normal(500), rng.standard_normal(500)
    het = np.abs(2 * f * (1 - f))
    o1, o2, bad = oracle.scale_effects(0, f, x1, x2, None, [-1.0], [0.5])
    assert bad == -1
    assert np.allclose(o1, x1 * np.sqrt(het ** -1.0), rtol=1e-15) and np.allclose(o2, x2 * np.sqrt(het ** 0.5), rtol=1e-15)
    o1, _, _ = oracle.scale_effects(1, f, x1)
    assert np.array_equal(o1, x1 / np.sqrt(het))
    o1, _, _ = oracle.scale_effects(2, f, x1)
    assert np.array_equal(o1, x1 * np.sqrt(het))
    f[7] = 0.0
    assert oracle.scale_effects(1, f, x1)[2] == 7 and oracle.scale_effects(2, f, x1)[2] == -1


def test_find_pheno_thread_split():
    """The oracle's sample loop split over threads / restricted to a column block keeps every sample's
    ascending-SNP operation order: bit-identical to the single-threaded restatement of source.cc:986-996
    (the full-size GPU parity tests rely on this)."""
    n, m = 10_007, 257
    rows = oracle.synth_rows(20240901, 3, m, n)
    _, f = oracle.find_freq(rows, n)
    rng = np.random.default_rng(0)
    x1, x2 = rng.standard_normal(m), rng.standard_normal(m)
    a1, a2 = oracle.find_pheno(rows, n, f, x1, x2)
    for t in (2, 7, 16):
        b1, b2 = oracle.find_pheno(rows, n, f, x1, x2, threads=t)
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
    c1, _ = oracle.find_pheno(rows, n, f, x1, None, samples=(4000, 6001))
    assert np.array_equal(c1[4000:6001], a1[4000:6001]) and not c1[:4000].any() and not c1[6001:].any()
    # a column block generated on its own equals the same columns of the full rows, and gives the same sums
    sub = oracle.synth_cols(20240901, 4000, 6048, first_row=3, n_rows=m)
    assert np.array_equal(sub, rows[:, 1000:1512])
    d1, d2 = oracle.find_pheno(sub, 2048, f, x1, x2, threads=3)
    assert np.array_equal(d1, a1[4000:6048]) and np.array_equal(d2, a2[4000:6048])
    ct, ft = oracle.find_freq_threaded(rows, n, 5)
    c0, f0 = oracle.find_freq(rows, n)
    assert np.array_equal(ct, c0) and np.array_equal(ft, f0)
