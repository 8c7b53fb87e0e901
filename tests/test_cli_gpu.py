"""The drop-in CLI end to end on a B200 against a Python restatement of the
reference pipeline (/root/reference/src/source.cc:1222-1316) built from the CPU
oracle (hot path) and tests/pyrng.py (draw order of SURVEY.md 9.2).

.pheno / .causals are text at 6 significant digits (source.cc:1066, :1101-1103),
so numbers are compared after parsing, at 2e-5 relative (1e-5 for FAST mode + print
rounding); case/control labels must agree exactly away from threshold ties.
"""
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

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "simu_b200", "bin", "simu_b200")
F32 = np.float32


def run(args, cwd, mode="fast", devices=None, extra_env=None):
    env = dict(os.environ, SIMU_B200_MODE=mode)
    if devices:
        env["SIMU_B200_DEVICES"] = devices
    env.update(extra_env or {})
    p = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def read_table(path):
    rows = [l.rstrip("\n").split("\t") for l in open(path)]
    return rows[0], rows[1:]


def close(a, b, rtol=1e-5):
    """|a-b| <= 1e-5 * max|b| (the FAST-mode bound, relative to the vector's scale)
    + 6-significant-digit print rounding of each number."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.all(np.abs(a - b) <= rtol * np.max(np.abs(b)) + 6e-6 * np.abs(b) + 1e-12)


def expected_pipeline(rows, n, m, seed, num_traits=1, causal_n=50, hsq=(0.7, 0.7), rg=0.0, gcta=False, cc=False, k=0.1):
    """source.cc:1222-1306 with the oracle as the hot path."""
    rng = pyrng.Mt19937(seed)
    perm = pyrng.random_shuffle(list(range(m)), rng)                     # :829-831
    comp = np.full(m, -1)
    comp[perm[:causal_n]] = 0                                            # :833-838
    causal = np.flatnonzero(comp != -1)
    _, freq = oracle.find_freq(rows, n, row_index=causal)                # :1235
    normal = pyrng.NormalDraws(rng)                                      # :894 / :937
    x1 = np.zeros(causal.size); x2 = np.zeros(causal.size)
    for j in range(causal.size):                                         # ascending variant index
        if num_traits == 1:
            x1[j] = math.sqrt(float(F32(1.0))) * normal()
        else:
            a, b = normal(), normal()
            r = float(F32(rg))
            x1[j] = a
            x2[j] = r * a + math.sqrt(1 - r * r) * b
    if gcta:                                                             # :1245-1264
        het = np.abs(2 * freq * (1 - freq))
        x1 *= np.sqrt(het ** -1.0)
        x2 *= np.sqrt(het ** -1.0)
    y1, y2 = oracle.find_pheno(rows, n, freq, x1, x2 if num_traits == 2 else None, row_index=causal)   # :1295
    ys, xs = [y1, y2][:num_traits], [x1, x2][:num_traits]
    for t in range(num_traits):                                          # :1299-1300 (noise: N draws per trait)
        normal = pyrng.NormalDraws(rng)
        noise = np.array([normal() for _ in range(n)])
        ys[t], xs[t], _, _ = oracle.apply_heritability(float(F32(hsq[t])), ys[t], noise, xs[t])
    if cc:                                                               # :1303-1306
        for t in range(num_traits):
            idx = oracle.liability_order(ys[t])
            max_cas, max_con = oracle.liability_split(float(F32(k)), n)
            con = pyrng.random_shuffle(list(range(max_con)), rng)
            cas = pyrng.random_shuffle(list(range(max_con, n)), rng)
            ys[t] = oracle.liability_labels(float(F32(k)), max_cas, n - max_cas, idx, con, cas)
    return causal, freq, xs, ys


@pytest.fixture(scope="module")
def fileset(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    n, m = 1203, 900
    rows = oracle.synth_rows(20240901, 0, m, n)
    plink.write_fileset(str(d / "syn"), rows, n)
    return d, rows, n, m


@pytest.mark.parametrize("mode", ["fast", "exact"])
def test_qt_single_trait(fileset, mode):
    d, rows, n, m = fileset
    out = run(["--bfile", "syn", "--qt", "--causal-n", "50", "--hsq", "0.5", "--seed", "1", "--out", f"qt_{mode}"], d, mode)
    assert "Calculate allele frequencies (for causal markers only)..." in out and "Calculate phenotypes..." in out
    causal, freq, xs, ys = expected_pipeline(rows, n, m, 1, causal_n=50, hsq=(0.5,))
    hdr, body = read_table(d / f"qt_{mode}.pheno")
    assert hdr == ["FID", "IID", "trait1"] and len(body) == n
    assert body[0][:2] == ["id1_0", "id2_0"] and body[-1][:2] == [f"id1_{n-1}", f"id2_{n-1}"]
    assert close([r[2] for r in body], ys[0])
    hdr, body = read_table(d / f"qt_{mode}.1.causals")
    assert hdr == ["SNP", "CHR", "POS", "A1", "A2", "FRQ", "BETA_c1"]
    assert [r[0] for r in body] == [f"rs{v}" for v in causal]
    assert [r[1:5] for r in body[:1]] == [["1", str(causal[0] + 1), "A", "G"]]
    assert close([r[5] for r in body], freq) and close([r[6] for r in body], xs[0])


def test_two_traits_gcta_norm_effect(fileset):
    d, rows, n, m = fileset
    run(["--bfile", "syn", "--qt", "--num-traits", "2", "--rg", "0.5", "--gcta-sigma", "--norm-effect", "--causal-n", "120",
         "--hsq", "0.5", "0.3", "--seed", "7", "--out", "biv"], d)
    causal, freq, xs, ys = expected_pipeline(rows, n, m, 7, num_traits=2, causal_n=120, hsq=(0.5, 0.3), rg=0.5, gcta=True)
    hdr, body = read_table(d / "biv.pheno")
    assert hdr == ["FID", "IID", "trait1", "trait2"]
    assert close([r[2] for r in body], ys[0]) and close([r[3] for r in body], ys[1])
    het = np.abs(2 * freq * (1 - freq))
    for t in (0, 1):                                                     # --norm-effect: beta * sqrt(het) (source.cc:1090-1094)
        _, body = read_table(d / f"biv.{t+1}.causals")
        assert close([r[6] for r in body], xs[t] * np.sqrt(het))


def test_case_control_labels(fileset):
    d, rows, n, m = fileset
    run(["--bfile", "syn", "--cc", "--k", "0.1", "--causal-n", "80", "--seed", "3", "--out", "cc"], d, "exact")
    _, _, _, ys = expected_pipeline(rows, n, m, 3, causal_n=80, hsq=(0.7,), cc=True, k=0.1)
    _, body = read_table(d / "cc.pheno")
    got = np.array([float(r[2]) for r in body])
    assert set(np.unique(got)) <= {1.0, 2.0}
    assert np.count_nonzero(got == 2) == int(math.floor(float(F32(0.1) * F32(n))))
    assert np.array_equal(got, ys[0])                                    # no ties at the threshold in this panel


def test_causal_variants_file_with_effects(fileset):
    """--causal-variants with effect sizes, --hsq 1 (env_coef = 0: no dependence on the noise draws)."""
    d, rows, n, m = fileset
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(m, 40, replace=False))
    b1, b2 = rng.standard_normal(40), rng.standard_normal(40)
    with open(d / "causal.txt", "w") as f:
        for v, a, b in zip(sel, b1, b2):
            f.write(f"rs{v}\t{float(a)!r}\t{float(b)!r}\n")
    out = run(["--bfile", "syn", "--qt", "--num-traits", "2", "--causal-variants", "causal.txt", "--hsq", "1", "1", "--seed", "9",
               "--out", "cv"], d, "exact")
    assert "Use effect sizes provided in --causal-variants file." in out
    _, freq = oracle.find_freq(rows, n, row_index=sel)
    y1, y2 = oracle.find_pheno(rows, n, freq, b1, b2, row_index=sel)
    e1, _, g1, _ = oracle.apply_heritability(1.0, y1, np.zeros(n), None)
    e2, _, g2, _ = oracle.apply_heritability(1.0, y2, np.zeros(n), None)
    _, body = read_table(d / "cv.pheno")
    assert close([r[2] for r in body], e1) and close([r[3] for r in body], e2)
    _, body = read_table(d / "cv.2.causals")
    assert [r[0] for r in body] == [f"rs{v}" for v in sel] and close([r[6] for r in body], b2 * g2)


def test_bfile_chr_concatenation(tmp_path):
    n = 301
    sizes = {"1": 40, "2": 25, "3": 35}
    rows = oracle.synth_rows(20240901, 0, sum(sizes.values()), n)
    off = 0
    for lab, sz in sizes.items():
        plink.write_fileset(str(tmp_path / f"c{lab}"), rows[off:off + sz], n, chromosome=int(lab), first_index=off)
        off += sz
    out = run(["--bfile-chr", "c@", "--chr-labels", "1", "2", "3", "--qt", "--causal-pi", "0.3", "--seed", "11", "--out", "chr"],
              tmp_path, "exact")
    assert "Found 100 variants in total." in out
    causal, freq, xs, ys = expected_pipeline(rows, n, 100, 11, causal_n=int(F32(0.3) * 100), hsq=(0.7,))
    _, body = read_table(tmp_path / "chr.pheno")
    assert close([r[2] for r in body], ys[0])
    _, body = read_table(tmp_path / "chr.1.causals")
    chrom = lambda v: "1" if v < 40 else ("2" if v < 65 else "3")     # noqa: E731
    assert [(r[0], r[1]) for r in body] == [(f"rs{v}", chrom(v)) for v in causal]


@pytest.mark.parametrize("devices", ["0,0,0", "all"])
def test_snp_sharding_over_several_contexts(tmp_path, devices):
    """SIMU_B200_DEVICES: the causal list cut into contiguous slices, one context (and host thread) per
    slice, frequencies concatenated, partial phenotypes summed -- here three contexts on GPU 0 (runs on
    a one-GPU box) and every GPU of the box; slices cross the bfile boundaries of --bfile-chr."""
    n = 701
    sizes = {"1": 40, "2": 25, "3": 35}
    rows = oracle.synth_rows(20240901, 0, sum(sizes.values()), n)
    off = 0
    for lab, sz in sizes.items():
        plink.write_fileset(str(tmp_path / f"c{lab}"), rows[off:off + sz], n, chromosome=int(lab), first_index=off)
        off += sz
    args = ["--bfile-chr", "c@", "--chr-labels", "1", "2", "3", "--qt", "--num-traits", "2", "--rg", "0.3", "--gcta-sigma",
            "--causal-pi", "0.47", "--hsq", "0.6", "0.4", "--seed", "5"]
    for mode in ("exact", "fast"):
        run(args + ["--out", f"one_{mode}"], tmp_path, mode)
        run(args + ["--out", f"many_{mode}"], tmp_path, mode, devices=devices)
        for ext in (".pheno", ".1.causals", ".2.causals"):
            h1, b1 = read_table(tmp_path / f"one_{mode}{ext}")
            h2, b2 = read_table(tmp_path / f"many_{mode}{ext}")
            assert h1 == h2 and len(b1) == len(b2)
            ncol = len(h1)
            first_num = 2 if ext == ".pheno" else 5
            assert [r[:first_num] for r in b1] == [r[:first_num] for r in b2]
            for c in range(first_num, ncol):
                assert close([r[c] for r in b2], [r[c] for r in b1], rtol=1e-5 if mode == "fast" else 1e-12)
    # more contexts than causal variants: empty slices are fine
    run(["--bfile", "c1", "--qt", "--causal-n", "2", "--seed", "1", "--out", "tiny"], tmp_path, "fast", devices="0,0,0")


def test_causal_rows_larger_than_hbm_run_in_batches(tmp_path):
    """A causal set that does not fit the GPU (here: SIMU_B200_MAX_PANEL_MB=1 against 3.5 MB of causal rows, three
    bfiles, slices and batches crossing file boundaries) is staged panel by panel for each of the two passes, as
    the reference streams it from disk (source.cc:1235, :1295).  EXACT mode continues every sample's chain of FP64
    additions across batches: the outputs are byte-identical to the resident run.  FAST mode adds per-batch sums."""
    n = 30_011
    sizes = {"1": 170, "2": 90, "3": 211}
    rows = oracle.synth_rows(20240901, 0, sum(sizes.values()), n)
    off = 0
    for lab, sz in sizes.items():
        plink.write_fileset(str(tmp_path / f"c{lab}"), rows[off:off + sz], n, chromosome=int(lab), first_index=off)
        off += sz
    args = ["--bfile-chr", "c@", "--chr-labels", "1", "2", "3", "--qt", "--num-traits", "2", "--rg", "0.3", "--gcta-sigma",
            "--causal-pi", "1.0", "--hsq", "0.6", "0.4", "--seed", "5"]
    small = {"SIMU_B200_MAX_PANEL_MB": "1"}                         # 16-row groups of 120 KB: 8 groups = 128 rows per batch
    for mode in ("exact", "fast"):
        run(args + ["--out", f"res_{mode}"], tmp_path, mode)
        run(args + ["--out", f"bat_{mode}"], tmp_path, mode, extra_env=small)
        run(args + ["--out", f"bat3_{mode}"], tmp_path, mode, devices="0,0,0", extra_env=small)
        for ext in (".pheno", ".1.causals", ".2.causals"):
            a = open(tmp_path / f"res_{mode}{ext}").read()
            b = open(tmp_path / f"bat_{mode}{ext}").read()
            if mode == "exact":
                assert a == b, ext                                   # bit-identical sums -> identical text
            h1, b1 = read_table(tmp_path / f"res_{mode}{ext}")
            for other in ("bat", "bat3"):
                h2, b2 = read_table(tmp_path / f"{other}_{mode}{ext}")
                first_num = 2 if ext == ".pheno" else 5
                assert h1 == h2 and [r[:first_num] for r in b1] == [r[:first_num] for r in b2]
                for c in range(first_num, len(h1)):
                    assert close([r[c] for r in b2], [r[c] for r in b1], rtol=1e-9)


def test_unknown_mode_is_an_error_and_default_is_exact(fileset):
    d, rows, n, m = fileset
    args = ["--bfile", "syn", "--qt", "--causal-n", "50", "--hsq", "0.5", "--seed", "1"]
    env = {k: v for k, v in os.environ.items() if k != "SIMU_B200_MODE"}
    p = subprocess.run([CLI] + args + ["--out", "dflt"], cwd=d, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "simu_b200: find_pheno mode exact, 1 GPU" in p.stdout
    assert "simu_b200: find_pheno mode exact" in open(d / "dflt.log").read()
    run(args + ["--out", "expl"], d, "exact")
    assert open(d / "dflt.pheno").read() == open(d / "expl.pheno").read()
    for bad in ("EXACT", "fsat", "1"):
        p = subprocess.run([CLI] + args + ["--out", "bad"], cwd=d, capture_output=True, text=True, timeout=600,
                           env=dict(env, SIMU_B200_MODE=bad))
        assert p.returncode == 1 and f"SIMU_B200_MODE must be 'exact' or 'fast', got '{bad}'" in p.stdout


def test_v099_bed_header_is_accepted_and_sample_major_refused(fileset, tmp_path):
    """libplinkio/src/bed_header.c:56-102: a v0.99 file starts with ONE order byte (01 = one locus per row) and its data
    at offset 1; the reference runs it.  Order byte 00 (one sample per row), v1.00 or v0.99, is refused with the
    reference's text (source.cc:196-200)."""
    d, rows, n, m = fileset
    B = (n + 3) // 4
    body = rows[:, :B].tobytes()
    for name, header in (("v100", bytes([0x6C, 0x1B, 0x01])), ("v099", bytes([0x01]))):
        for ext in (".bim", ".fam"):
            (tmp_path / f"{name}{ext}").write_bytes((d / f"syn{ext}").read_bytes())
        (tmp_path / f"{name}.bed").write_bytes(header + body)
        run(["--bfile", name, "--qt", "--causal-n", "60", "--seed", "3", "--out", name], tmp_path, "exact")
    assert open(tmp_path / "v100.pheno").read() == open(tmp_path / "v099.pheno").read()
    assert open(tmp_path / "v100.1.causals").read() == open(tmp_path / "v099.1.causals").read()
    for name, header in (("s100", bytes([0x6C, 0x1B, 0x00])), ("s099", bytes([0x00])), ("pre", bytes([0x7F]))):
        for ext in (".bim", ".fam"):
            (tmp_path / f"{name}{ext}").write_bytes((d / f"syn{ext}").read_bytes())
        (tmp_path / f"{name}.bed").write_bytes(header + body)
        p = subprocess.run([CLI, "--bfile", name, "--qt", "--out", name], cwd=tmp_path, capture_output=True, text=True, timeout=600)
        assert p.returncode == 1 and f"ERROR: Unsupported format in {name}, snps must be rows, samples must be columns" in p.stdout


def test_causal_regions_and_components(fileset):
    """--causal-regions with two mixture components: one shuffle per region file
    (source.cc:816-826), effects drawn in ascending variant order with per-component sigsq."""
    d, rows, n, m = fileset
    reg = [list(range(10, 200)), list(range(400, 650))]
    for c, r in enumerate(reg):
        with open(d / f"region{c}.txt", "w") as f:
            f.write("\n".join(f"rs{v}" for v in r) + "\n")
    run(["--bfile", "syn", "--qt", "--num-components", "2", "--causal-n", "20", "30", "--causal-regions", "region0.txt",
         "region1.txt", "--trait1-sigsq", "1", "0.25", "--hsq", "0.6", "--seed", "21", "--out", "reg"], d, "exact")
    rng = pyrng.Mt19937(21)
    comp = np.full(m, -1)
    for c, (r, k) in enumerate(zip(reg, (20, 30))):
        idx = pyrng.random_shuffle(list(r), rng)
        comp[idx[:k]] = c
    causal = np.flatnonzero(comp != -1)
    _, freq = oracle.find_freq(rows, n, row_index=causal)
    normal = pyrng.NormalDraws(rng)
    sig = [math.sqrt(float(F32(1.0))), math.sqrt(float(F32(0.25)))]
    x = np.array([sig[comp[v]] * normal() for v in causal])
    y, _ = oracle.find_pheno(rows, n, freq, x, None, row_index=causal)
    normal = pyrng.NormalDraws(rng)
    noise = np.array([normal() for _ in range(n)])
    y, x, _, _ = oracle.apply_heritability(float(F32(0.6)), y, noise, x)
    _, body = read_table(d / "reg.pheno")
    assert close([r[2] for r in body], y)
    hdr, body = read_table(d / "reg.1.causals")
    assert hdr[-2:] == ["BETA_c1", "BETA_c2"]
    assert [r[0] for r in body] == [f"rs{v}" for v in causal]
    got = np.array([[float(r[6]), float(r[7])] for r in body])
    exp = np.zeros_like(got)
    exp[np.arange(causal.size), comp[causal]] = x
    assert close(got.ravel(), exp.ravel())


def test_trait2_snp_offset(fileset):
    """--trait2-snp-offset: three implicit components (source.cc:691-695, :731-757, :903-924)."""
    d, rows, n, m = fileset
    run(["--bfile", "syn", "--qt", "--num-traits", "2", "--causal-n", "40", "--trait2-snp-offset", "1", "--hsq", "0.5", "0.5",
         "--seed", "33", "--out", "off"], d, "exact")
    rng = pyrng.Mt19937(33)
    perm = pyrng.random_shuffle(list(range(m)), rng)
    comp = np.full(m, -1)
    for s1 in perm[:40]:
        comp[s1] += 1
        comp[(s1 + 1) % m] += 2
    causal = np.flatnonzero(comp != -1)
    _, freq = oracle.find_freq(rows, n, row_index=causal)
    normal = pyrng.NormalDraws(rng)
    x1 = np.zeros(causal.size); x2 = np.zeros(causal.size)
    for j, v in enumerate(causal):
        a, b = normal(), normal()
        if comp[v] in (0, 2):
            x1[j] = a
        if comp[v] in (1, 2):
            x2[j] = b
    y1, y2 = oracle.find_pheno(rows, n, freq, x1, x2, row_index=causal)
    out = []
    for y, x in ((y1, x1), (y2, x2)):
        normal = pyrng.NormalDraws(rng)
        noise = np.array([normal() for _ in range(n)])
        out.append(oracle.apply_heritability(float(F32(0.5)), y, noise, x))
    _, body = read_table(d / "off.pheno")
    assert close([r[2] for r in body], out[0][0]) and close([r[3] for r in body], out[1][0])
    hdr, body = read_table(d / "off.2.causals")
    assert hdr[-3:] == ["BETA_c1", "BETA_c2", "BETA_c3"] and len(body) == causal.size


def test_config1_end_to_end(tmp_path):
    """BASELINE configs[0] through the drop-in CLI: 10K individuals x 100K SNPs on disk (250 MB .bed),
    `--qt --causal-n 1000 --hsq 0.5 --seed 1`, against the oracle pipeline on the same files."""
    import time
    from simu_b200 import capi
    n, m = 10_000, 100_000
    with capi.Context(n) as ctx:                      # write the panel with the device generator
        ctx.panel_alloc(m)
        ctx.panel_synth(20240901, 0, m)
        rows = ctx.panel_get_rows(0, m)
    plink.write_fileset(str(tmp_path / "syn"), rows, n)
    t0 = time.perf_counter()
    out = run(["--bfile", "syn", "--qt", "--causal-n", "1000", "--hsq", "0.5", "--seed", "1", "--out", "c1"], tmp_path, "fast")
    wall = time.perf_counter() - t0
    assert "done, 10000 samples, 100000 variants." in out
    causal, freq, xs, ys = expected_pipeline(rows, n, m, 1, causal_n=1000, hsq=(0.5,))
    _, body = read_table(tmp_path / "c1.pheno")
    assert len(body) == n and close([r[2] for r in body], ys[0])
    _, body = read_table(tmp_path / "c1.1.causals")
    assert [r[0] for r in body] == [f"rs{v}" for v in causal]
    assert close([r[5] for r in body], freq) and close([r[6] for r in body], xs[0])
    print(f"config1 CLI wall time {wall:.2f} s")
    assert wall < 60


# ------------------------------------------------------------------------------------------
# product CLI against the REFERENCE CLI itself (unmodified source.cc + oracle/boost_shim)

def _run_ref(args, cwd):
    p = subprocess.run([oracle.SIMU_REF_BIN] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


REF_CASES = [
    ["--qt", "--causal-n", "50", "--hsq", "0.5", "--seed", "1"],
    ["--qt", "--num-traits", "2", "--rg", "-0.5", "--gcta-sigma", "--norm-effect", "--causal-pi", "0.1", "--hsq", "0.5", "0.3", "--seed", "7"],
    ["--cc", "--k", "0.1", "--causal-n", "80", "--seed", "3"],
    ["--cc", "--num-traits", "2", "--k", "0.2", "0.1", "--ncas", "100", "60", "--ncon", "200", "300", "--causal-n", "40", "--seed", "5"],
    ["--qt", "--num-components", "2", "--causal-n", "20", "30", "--trait1-sigsq", "1", "0.25", "--trait1-s-pow", "-0.25", "-0.5",
     "--seed", "9"],
    ["--qt", "--num-traits", "2", "--causal-n", "40", "--trait2-snp-offset", "1", "--seed", "33"],
]


@pytest.mark.skipif(not os.path.exists(oracle.SIMU_REF_BIN), reason="oracle/_ref/simu_ref absent")
@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("case", range(len(REF_CASES)))
def test_product_cli_equals_reference_cli(fileset, case, mode):
    d, rows, n, m = fileset
    args = ["--bfile", "syn"] + REF_CASES[case]
    ref_log = _run_ref(args + ["--out", f"ref{case}"], d)
    our_log = run(args + ["--out", f"our{case}_{mode}"], d, mode)
    k = 2 if "2" == (args[args.index("--num-traits") + 1] if "--num-traits" in args else "1") else 1
    rh, rb = read_table(d / f"ref{case}.pheno")
    oh, ob = read_table(d / f"our{case}_{mode}.pheno")
    assert rh == oh and [r[:2] for r in rb] == [r[:2] for r in ob]
    for t in range(k):
        ref_y = np.array([float(r[2 + t]) for r in rb]); our_y = np.array([float(r[2 + t]) for r in ob])
        if "--cc" in args:
            assert np.array_equal(ref_y, our_y)                    # labels: exact
        else:
            assert close(our_y, ref_y)
        rh2, rb2 = read_table(d / f"ref{case}.{t+1}.causals")
        oh2, ob2 = read_table(d / f"our{case}_{mode}.{t+1}.causals")
        assert rh2 == oh2 and [r[:6] for r in rb2] == [r[:6] for r in ob2]       # SNP CHR POS A1 A2 FRQ: text-identical
        assert close(np.array([[float(v) for v in r[6:]] for r in ob2]).ravel(),
                     np.array([[float(v) for v in r[6:]] for r in rb2]).ravel())
    # the log has the same stage lines (banner version line and timestamps differ)
    stage = lambda s: [l for l in s.split("\n") if l.endswith("... ") or l.endswith("...") or l.startswith("\t--")]   # noqa: E731
    assert stage(ref_log) == [l.replace(f"our{case}_{mode}", f"ref{case}") for l in stage(our_log)]
