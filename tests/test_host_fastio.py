"""Host-side text I/O of the drop-in CLI: the buffered writers and the
in-place .bim/.fam tokeniser must produce exactly what the iostream code they replace produces
(which tests/test_cli_gpu.py pins against the reference binary's .pheno / .causals)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(src, out):
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-pthread", src, "-o", out], check=True, cwd=ROOT)


def test_textwriter_formats_like_ostream(tmp_path):
    exe = str(tmp_path / "fmt_check")
    _build("tests/helpers/fmt_check.cc", exe)
    r = subprocess.run([exe, str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 differences" in r.stdout


def test_sharded_name_index_reports_duplicates_like_the_sequential_build(tmp_path):
    exe = str(tmp_path / "name_index_check")
    _build("tests/helpers/name_index_check.cc", exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_threaded_bim_parse_equals_single_thread(tmp_path):
    exe = str(tmp_path / "bim_parse_check")
    _build("tests/helpers/bim_parse_check.cc", exe)
    r = subprocess.run([exe, str(tmp_path / "big.bim")], capture_output=True, text=True)
    assert r.returncode == 0 and "threaded == single: yes" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("m,n", [(1, 1), (50_000, 3_000), (300_000, 100)])      # the last one takes the threaded paths
def test_parser_map_and_writers_match_iostream_restatement(tmp_path, m, n):
    exe = str(tmp_path / "host_io_bench")
    _build("tools/host_io_bench.cc", exe)
    r = subprocess.run([exe, str(m), str(n), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout)
    assert out["identical"] is True and out["n_variants"] == m
