"""The C-ABI library loads, exports every symbol include/simu_b200.h declares,
and fails loudly (no CPU fallback) when there is no GPU.  CPU only."""
import ctypes as C
import os
import re

import pytest

from simu_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "simu_b200.h")).read()
    return sorted(set(re.findall(r"SIMU_B200_API[^;]*?\b(simu_b200_\w+)\s*\(", text)))


def test_library_is_built():
    assert os.path.exists(capi.LIB_PATH), "run `python -m simu_b200.build`"


def test_exports_every_declared_symbol():
    lib = C.CDLL(capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/simu_b200.h but not exported"


def test_binding_covers_header():
    assert sorted(capi.SIGNATURES) == _declared()
    capi.load()


def test_row_bytes_matches_bed_header():
    lib = capi.load()
    assert [lib.simu_b200_row_bytes(n) for n in (1, 4, 5, 100000, 500000)] == [1, 1, 2, 25000, 125000]


def test_no_cpu_fallback(has_gpu):
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(capi.SimuB200Error) as ei:
        capi.Context(100)
    assert ei.value.status == capi.ECUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "simu_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("SURVEY", ""), f"{f} mentions the oracle"
