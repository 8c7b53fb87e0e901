"""Pin the oracle and the host pipeline against the reference ITSELF: the unmodified
/root/reference/src/source.cc compiled against oracle/boost_shim/ (oracle/Makefile ->
oracle/_ref/libsimu_refsrc.so and oracle/_ref/simu_ref).  The hot-path functions contain no
Boost, so these are the literal reference loops; RNG and option parsing go through the
stand-in headers (same published algorithms as simu_b200/host/rng.h, not pinned against a real
Boost build).  CPU only."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pyrng  # noqa: E402
from oracle import oracle  # noqa: E402
from simu_b200 import plink  # noqa: E402

pytestmark = pytest.mark.skipif(oracle.refsrc() is None or not os.path.exists(oracle.SIMU_REF_BIN),
                                reason="oracle/_ref reference-source build absent")
F32 = np.float32


@pytest.fixture(scope="module")
def fileset(tmp_path_factory):
    d = tmp_path_factory.mktemp("refsrc")
    n, m = 403, 300
    rows = oracle.synth_rows(20240901, 0, m, n)
    rows[7, :] = 0x55                                   # an all-missing SNP
    plink.write_fileset(str(d / "syn"), rows, n)
    return d, rows, n, m


def test_oracle_find_freq_pheno_equal_reference_source(fileset):
    d, rows, n, m = fileset
    rng = np.random.default_rng(3)
    comp = np.where(rng.random(m) < 0.4, 0, -1).astype(np.int32)
    comp[[0, 7, m - 1]] = 0
    causal = np.flatnonzero(comp != -1)
    x1, x2 = rng.standard_normal(m), rng.standard_normal(m)
    for k2 in (False, True):
        freq, p1, p2, _ = oracle.refsrc_freq_pheno(str(d / "syn"), comp, x1, x2 if k2 else None)
        oc, of = oracle.find_freq(rows, n, row_index=causal)
        assert np.array_equal(of, freq[causal]) and np.all(np.isnan(freq[comp == -1]))
        o1, o2 = oracle.find_pheno(rows, n, of, x1[causal], x2[causal] if k2 else None, row_index=causal)
        assert np.array_equal(o1, p1)
        if k2:
            assert np.array_equal(o2, p2)


def test_oracle_apply_heritability_equals_reference_source():
    rng = np.random.default_rng(5)
    y = rng.standard_normal(999) * 3
    eff = rng.standard_normal(40)
    for hsq in (0.5, 0.7, 0.0, 1.0):
        rp, re, noise = oracle.refsrc_apply_heritability(float(F32(hsq)), 77, y, eff)
        op, oe, _, _ = oracle.apply_heritability(float(F32(hsq)), y, noise, eff)
        assert np.array_equal(rp, op) and np.array_equal(re, oe)
    # and the noise is what the Python restatement of the RNG layer draws
    e = pyrng.Mt19937(77)
    nd = pyrng.NormalDraws(e)
    assert np.array_equal(noise, np.array([nd() for _ in range(999)]))


def test_oracle_liability_threshold_equals_reference_source():
    rng = np.random.default_rng(6)
    y = rng.standard_normal(500)
    k, seed = float(F32(0.2)), 123
    got = oracle.refsrc_apply_liability_threshold(k, 60, 150, seed, y)
    e = pyrng.Mt19937(seed)
    max_cas, max_con = oracle.liability_split(k, y.size)
    con = pyrng.random_shuffle(list(range(max_con)), e)
    cas = pyrng.random_shuffle(list(range(max_con, y.size)), e)
    exp = oracle.liability_labels(k, 60, 150, oracle.liability_order(y), con, cas)
    assert np.array_equal(got, exp)
    assert (np.count_nonzero(got == 2), np.count_nonzero(got == 1), np.count_nonzero(got == -9)) == (60, 150, 290)


def _run_ref(args, cwd):
    p = subprocess.run([oracle.SIMU_REF_BIN] + args, cwd=cwd, capture_output=True, text=True, timeout=300)
    return p.returncode, p.stdout, p.stderr


def _table(path):
    rows = [l.rstrip("\n").split("\t") for l in open(path)]
    return rows[0], rows[1:]


def test_reference_cli_equals_python_pipeline(fileset):
    """The whole reference main() (source.cc:1111-1344) against the oracle + pyrng pipeline:
    draw order of SURVEY.md 9.2, gcta scaling, heritability, output formats."""
    d, rows, n, m = fileset
    rc, out, err = _run_ref(["--bfile", "syn", "--qt", "--num-traits", "2", "--rg", "0.5", "--gcta-sigma", "--norm-effect",
                             "--causal-n", "50", "--hsq", "0.5", "0.3", "--seed", "7", "--out", "ref"], d)
    assert rc == 0, out + err
    rng = pyrng.Mt19937(7)
    perm = pyrng.random_shuffle(list(range(m)), rng)
    comp = np.full(m, -1); comp[perm[:50]] = 0
    causal = np.flatnonzero(comp != -1)
    _, freq = oracle.find_freq(rows, n, row_index=causal)
    nd = pyrng.NormalDraws(rng)
    x1 = np.zeros(50); x2 = np.zeros(50)
    r = float(F32(0.5))
    for j in range(50):
        a, b = nd(), nd()
        x1[j], x2[j] = a, r * a + math.sqrt(1 - r * r) * b
    het = np.abs(2 * freq * (1 - freq))
    x1 *= np.sqrt(het ** -1.0); x2 *= np.sqrt(het ** -1.0)
    y1, y2 = oracle.find_pheno(rows, n, freq, x1, x2, row_index=causal)
    outs = []
    for y, x, h in ((y1, x1, 0.5), (y2, x2, 0.3)):
        nd = pyrng.NormalDraws(rng)
        noise = np.array([nd() for _ in range(n)])
        outs.append(oracle.apply_heritability(float(F32(h)), y, noise, x))
    hdr, body = _table(d / "ref.pheno")
    assert hdr == ["FID", "IID", "trait1", "trait2"]
    fmt = lambda v: "%g" % v                                   # default ostream precision = %g     # noqa: E731
    assert [r_[2] for r_ in body] == [fmt(v) for v in outs[0][0]]
    assert [r_[3] for r_ in body] == [fmt(v) for v in outs[1][0]]
    for t in (0, 1):
        hdr, body = _table(d / f"ref.{t+1}.causals")
        assert [r_[0] for r_ in body] == [f"rs{v}" for v in causal]
        assert [r_[5] for r_ in body] == [fmt(v) for v in freq]
        assert [r_[6] for r_ in body] == [fmt(v) for v in outs[t][1] * np.sqrt(het)]


def test_reference_cli_case_control_and_product_cli_messages(fileset, tmp_path):
    d, rows, n, m = fileset
    rc, out, _ = _run_ref(["--bfile", "syn", "--cc", "--k", "0.1", "--causal-n", "30", "--seed", "3", "--out", "cc"], d)
    assert rc == 0
    _, body = _table(d / "cc.pheno")
    lab = np.array([float(r_[2]) for r_ in body])
    assert np.count_nonzero(lab == 2) == int(math.floor(float(F32(0.1) * F32(n)))) and set(np.unique(lab)) <= {1.0, 2.0}
    # the product CLI must produce the same validation / parse messages as the reference main()
    cli = os.path.join(ROOT, "simu_b200", "bin", "simu_b200")
    if not os.path.exists(cli):
        pytest.skip("product CLI not built")
    for args in (["--qt"], ["--bfile", "x"], ["--bfile", "x", "--qt", "--num-traits", "3"], ["--bfile", "nope", "--qt"],
                 ["--bfile", "x", "--qt", "--hsq", "0.2", "0.3"], ["--bfile", "x", "--cc", "--ncas", "5"],
                 ["--bfile", "x", "--qt", "--num-traits", "abc"], ["--bfile", "x", "--frobnicate"], ["--bfile"],
                 ["--bfile", "a", "--bfile", "b"], ["--bf", "x"],
                 ["--bfile", "x", "--qt", "--causal-pi", "0.1", "--causal-n", "5"]):
        r1 = subprocess.run([oracle.SIMU_REF_BIN] + args, cwd=tmp_path, capture_output=True, text=True)
        r2 = subprocess.run([cli] + args, cwd=tmp_path, capture_output=True, text=True)
        last = lambda s: [l for l in s.strip().split("\n") if l][-1] if s.strip() else ""     # noqa: E731
        assert r1.returncode == r2.returncode == 1
        assert last(r1.stdout + r1.stderr) == last(r2.stdout + r2.stderr), args


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,k2", [(403, 300, True), (1001, 257, False), (4099, 130, True)])
def test_cuda_path_equals_reference_source(tmp_path, n, m, k2):
    """The CUDA path through the C ABI against the LITERAL reference functions on the same bfile:
    allele frequencies bit-exact, EXACT-mode phenotypes bit-exact, FAST mode within 1e-5,
    apply_heritability within 1e-12 with the reference's own noise draws."""
    from simu_b200 import capi
    rows = oracle.synth_rows(20240901, 11, m, n)
    rows[3, :] = 0x55
    prefix = str(tmp_path / "syn")
    plink.write_fileset(prefix, rows, n)
    rng = np.random.default_rng(n)
    comp = np.where(rng.random(m) < 0.5, 0, -1).astype(np.int32)
    comp[[0, 3, m - 1]] = 0
    causal = np.flatnonzero(comp != -1)
    x1, x2 = rng.standard_normal(m), rng.standard_normal(m)
    freq, p1, p2, _ = oracle.refsrc_freq_pheno(prefix, comp, x1, x2 if k2 else None)
    with capi.Context(n) as ctx:
        ctx.panel_alloc(causal.size)
        ctx.panel_load_bed(prefix + ".bed", m, causal)
        _, f = ctx.find_freq()
        assert np.array_equal(f, freq[causal])
        e1, e2 = ctx.find_pheno(f, x1[causal], x2[causal] if k2 else None, mode=capi.MODE_EXACT)
        assert np.array_equal(e1, p1) and (not k2 or np.array_equal(e2, p2))
        g1, g2 = ctx.find_pheno(f, x1[causal], x2[causal] if k2 else None, mode=capi.MODE_FAST)
        assert np.max(np.abs(g1 - p1)) <= 1e-5 * np.max(np.abs(p1))
        if k2:
            assert np.max(np.abs(g2 - p2)) <= 1e-5 * np.max(np.abs(p2))
        rp, re, noise = oracle.refsrc_apply_heritability(float(F32(0.6)), 99, p1, x1)
        gy, gx, _, _ = ctx.apply_heritability(float(F32(0.6)), e1, noise, x1)
        assert np.max(np.abs(gy - rp)) <= 1e-12 * np.max(np.abs(rp)) and np.max(np.abs(gx - re)) <= 1e-12 * np.max(np.abs(re))
        lab = oracle.refsrc_apply_liability_threshold(float(F32(0.1)), 10, 50, 5, rp)
        idx = ctx.liability_order(rp)
        e = pyrng.Mt19937(5)
        mc_, mn_ = oracle.liability_split(float(F32(0.1)), n)
        con = pyrng.random_shuffle(list(range(mn_)), e); cas = pyrng.random_shuffle(list(range(mn_, n)), e)
        assert np.array_equal(lab, oracle.liability_labels(float(F32(0.1)), 10, 50, idx, con, cas))
