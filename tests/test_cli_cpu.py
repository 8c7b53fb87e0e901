"""Host side of the drop-in CLI (C++), CPU only: option surface and validation
messages of /root/reference/src/source.cc:423-700 / :1116-1207, and the RNG
layer against an independent Python restatement of the same algorithms."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
CLI = os.path.join(ROOT, "simu_b200", "bin", "simu_b200")


@pytest.fixture(scope="module", autouse=True)
def _built():
    from simu_b200 import build
    build.build_library()
    build.build_cli()
    assert os.path.exists(CLI)


def run(args, cwd):
    p = subprocess.run([CLI] + args, cwd=cwd, capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout, p.stderr


@pytest.mark.parametrize("args,needle,where", [
    (["--qt"], "ERROR: Either --bfile or --bfile-chr must be specified", "out"),
    (["--bfile", "x"], "ERROR: Either --qt or --cc must be specified", "out"),
    (["--bfile", "x", "--qt", "--num-traits", "3"], "ERROR: --num-traits must be either 1 or 2", "out"),
    (["--bfile", "x", "--qt", "--num-components", "0"], "ERROR: --num-components must be non-negative number", "out"),
    (["--bfile", "x", "--cc", "--ncas", "5"], "ERROR: Options --ncas and --ncon must be either both present, or both absent", "out"),
    (["--bfile", "x", "--cc", "--k", "1.5"], "ERROR: Prevalence --k must be between 0.0 and 1.0", "out"),
    (["--bfile", "x", "--qt", "--hsq", "0.2", "0.3"], "ERROR: Number of --hsq values does not match the number of traits", "out"),
    (["--bfile", "x", "--qt", "--hsq", "-0.5"], "ERROR: Heritability --hsq must be between 0.0 and 1.0", "out"),
    (["--bfile", "x", "--qt", "--causal-pi", "0.1", "--causal-n", "5"],
     "ERROR: At most one of --causal-pi, --causal-n, --causal-variants options must be used", "out"),
    (["--bfile", "x", "--qt", "--num-traits", "2", "--rg", "-1.5"], "ERROR: --rg values must be between -1.0 and 1.0", "out"),
    (["--bfile", "x", "--qt", "--gcta-sigma", "--trait1-s-pow", "-0.25"],
     "ERROR: --gcta-sigma is incompatible with --trait1-s-pow / --trait2-s-pow", "out"),
    (["--bfile", "x", "--qt", "--causal-variants", "missing.txt"], "ERROR: input file missing.txt does not exist", "out"),
    (["--bfile", "nope", "--qt"], "ERROR: Could not open nope", "out"),
    (["--bfile", "x", "--qt", "--num-traits", "abc"], "Exception  : the argument ('abc') for option '--num-traits' is invalid", "err"),
    (["--bfile", "x", "--frobnicate"], "Exception  : unrecognised option '--frobnicate'", "err"),
    (["--bfile"], "Exception  : the required argument for option '--bfile' is missing", "err"),
    (["--bfile", "a", "--bfile", "b"], "Exception  : option '--bfile' cannot be specified more than once", "err"),
    (["stray", "12"], "ERROR: Either --bfile or --bfile-chr must be specified", "out"),   # nameless positionals are skipped
])
def test_validation_messages(tmp_path, args, needle, where):
    rc, out, err = run(args, tmp_path)
    assert rc == 1                                         # EXIT_FAILURE (source.cc:1333, :1337)
    assert needle in (out if where == "out" else err)


def test_help_goes_to_stderr_and_fails(tmp_path):
    rc, out, err = run(["--help"], tmp_path)               # source.cc:1176-1207
    assert rc == 1 and "--causal-variants arg" in err and "Examples:" in err and out == ""


def test_log_file_and_banner(tmp_path):
    rc, out, _ = run(["--bfile", "nope", "--qt", "--out", "res"], tmp_path)
    log = (tmp_path / "res.log").read_text()
    assert log == out and "* SIMU - library for simulation of GWAS summary statistics" in log
    assert "Call:" in log and "--bfile nope" in log.replace("\\\n\t", "")


def test_negative_numbers_are_values(tmp_path):
    # source.cc:157-185: "-0.5" must be taken as the value of --rg, not as an option
    rc, out, err = run(["--bfile", "nope", "--qt", "--num-traits", "2", "--rg", "-0.5"], tmp_path)
    assert "unrecognised" not in err and "ERROR: Could not open nope" in out


def test_rng_layer_matches_python_restatement(tmp_path):
    import pyrng
    exe = str(tmp_path / "rng_dump")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "helpers", "rng_dump.cc"), "-o", exe], check=True)
    for seed, n in ((1, 50), (123456789, 200), (2**40 + 17, 31), (-5, 10)):
        lines = dict(l.split(" ", 1) for l in subprocess.run([exe, str(seed), str(n)], capture_output=True, text=True,
                                                             check=True).stdout.strip().split("\n"))
        e = pyrng.Mt19937(seed)
        assert [int(x) for x in lines["shuffle"].split()] == pyrng.random_shuffle(list(range(n)), e)
        d = pyrng.NormalDraws(e)
        assert [float(x) for x in lines["normal"].split()] == [d() for _ in range(n)]
        d = pyrng.NormalDraws(e)
        assert [float(x) for x in lines["normal2"].split()] == [d() for _ in range(3)]
        assert [int(x) for x in lines["raw"].split()] == [e(), e()]


def test_mt19937_known_answer():
    """10000th output of mt19937 seeded with 5489 is 4123659995 (C++11 [rand.predef])."""
    import pyrng
    e = pyrng.Mt19937(5489)
    for _ in range(9999):
        e()
    assert e() == 4123659995
