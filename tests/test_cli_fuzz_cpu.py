"""Randomised command lines through the product CLI's host side and through the reference
main() (unmodified source.cc + oracle/boost_shim): option parsing, defaults, cross-option
validation (source.cc:423-700) and the "Options in effect" echo (:702-729) must agree line
for line.  CPU only: on a box without a GPU the product CLI stops right after that echo with
its "no CPU fallback" error, which is exactly the part compared here."""
import os
import random
import shutil
import subprocess

import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "simu_b200", "bin", "simu_b200")
G = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.skipif(not os.path.exists(oracle.SIMU_REF_BIN), reason="oracle/_ref/simu_ref absent")


def _cmdline(rng):
    a = ["--bfile", "syn_small"] if rng.random() < 0.9 else []
    a += rng.choice([["--qt"], ["--cc"], ["--qt", "--cc"], []])
    nt = rng.choice([1, 1, 2, 2, 3])
    if rng.random() < 0.8:
        a += ["--num-traits", str(nt)]
    nc = rng.choice([1, 1, 2, 3])
    if rng.random() < 0.5:
        a += ["--num-components", str(nc)]
    vals = lambda n, lo, hi: [f"{rng.uniform(lo, hi):.3g}" for _ in range(n)]            # noqa: E731
    cnt = lambda n: n if rng.random() < 0.8 else rng.choice([1, 2, 3])                       # noqa: E731
    if rng.random() < 0.5:
        a += ["--hsq"] + vals(cnt(min(nt, 2)), -0.1, 1.1)
    if rng.random() < 0.4:
        a += ["--k"] + vals(cnt(min(nt, 2)), -0.1, 1.1)
    if rng.random() < 0.2:
        a += ["--ncas"] + [str(rng.randint(0, 30)) for _ in range(cnt(min(nt, 2)))]
    if rng.random() < 0.2:
        a += ["--ncon"] + [str(rng.randint(0, 150)) for _ in range(cnt(min(nt, 2)))]
    which = rng.choice(["pi", "n", "none", "both"])
    if which in ("pi", "both"):
        a += ["--causal-pi"] + vals(cnt(nc), -0.05, 0.6)
    if which in ("n", "both"):
        a += ["--causal-n"] + [str(rng.randint(-1, 40)) for _ in range(cnt(nc))]
    if rng.random() < 0.3:
        a += ["--trait1-sigsq"] + vals(cnt(nc), -0.2, 2)
    if rng.random() < 0.3:
        a += ["--trait2-sigsq"] + vals(cnt(nc), -0.2, 2)
    if rng.random() < 0.4:
        a += ["--rg"] + vals(cnt(nc), -1.2, 1.2)
    if rng.random() < 0.2:
        a += ["--trait1-s-pow"] + vals(cnt(nc), -1, 0)
    if rng.random() < 0.15:
        a += ["--trait2-s-pow"] + vals(cnt(nc), -1, 0)
    if rng.random() < 0.2:
        a += ["--gcta-sigma"]
    if rng.random() < 0.2:
        a += ["--norm-effect"]
    if rng.random() < 0.15:
        a += ["--trait2-snp-offset", str(rng.randint(0, 3))]
    if rng.random() < 0.1:
        a += [rng.choice(["stray", "-7", "3.5"])]
    a += ["--seed", str(rng.randint(1, 1000))]
    rng.shuffle_groups = None
    return a


def _echo_block(text):
    """Everything from 'Opening' (or the first ERROR/WARNING) through the options echo."""
    lines = text.split("\n")
    keep, on = [], False
    for l in lines:
        if l.startswith(("Opening", "WARNING", "ERROR", "Found ", "Validate", "Reading")):
            on = True
        if on:
            if l.startswith(("Calculate allele", "ERROR: no CUDA", "Analysis finished")):
                break
            keep.append(l)
    return keep


@pytest.mark.parametrize("chunk", range(6))
def test_random_command_lines_agree_with_reference(tmp_path, chunk):
    for ext in (".bed", ".bim", ".fam"):
        shutil.copy(os.path.join(G, "syn_small" + ext), tmp_path / ("syn_small" + ext))
    rng = random.Random(1000 + chunk)
    for it in range(40):
        args = _cmdline(rng)
        r = subprocess.run([oracle.SIMU_REF_BIN] + args + ["--out", "ref"], cwd=tmp_path, capture_output=True, text=True)
        o = subprocess.run([CLI] + args + ["--out", "our"], cwd=tmp_path, capture_output=True, text=True)
        for f in os.listdir(tmp_path):                     # the reference run may have completed: drop its outputs
            if f.startswith(("ref.", "our.")):
                os.remove(tmp_path / f)
        rb = [l.replace("--out ref", "--out X") for l in _echo_block(r.stdout)]
        ob = [l.replace("--out our", "--out X") for l in _echo_block(o.stdout)]
        assert rb == ob, (args, rb, ob)
        assert r.stderr.replace("ref", "X") == o.stderr.replace("our", "X"), args     # "Exception  : ..." parse errors
        if r.returncode != 0:
            assert o.returncode != 0, args
