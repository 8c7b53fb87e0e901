"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; nothing under ``simu_b200/`` does.  It wraps

* ``liboracle.so``           -- oracle/simu_oracle.c, the restatement of
  /root/reference/src/source.cc:843-1055 (+ the libplinkio 2-bit decode);
* ``_ref/libsimu_ref.so``    -- the reference's own libplinkio compiled from
  /root/reference/libplinkio (oracle/Makefile) plus oracle/ref_harness.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    need = force or not os.path.exists(os.path.join(_HERE, "liboracle.so"))
    src_t = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("simu_oracle.c", "simu_oracle.h"))
    if not need and os.path.getmtime(os.path.join(_HERE, "liboracle.so")) < src_t:
        need = True
    if need or (os.path.isdir("/root/reference/libplinkio/src")
                and not all(os.path.exists(os.path.join(_HERE, "_ref", f))
                            for f in ("libsimu_ref.so", "libsimu_refsrc.so", "simu_ref"))):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.oracle_bed_header.argtypes = [_u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_bed_header.restype = C.c_int
        L.oracle_bed_row_bytes.argtypes = [C.c_size_t]
        L.oracle_bed_row_bytes.restype = C.c_size_t
        L.oracle_unpack_snps.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_pack_snps.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_find_freq.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                       C.c_void_p, C.c_void_p]
        L.oracle_find_pheno.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_find_pheno_range.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_apply_heritability.argtypes = [C.c_float, C.c_void_p, C.c_size_t, C.c_void_p,
                                                C.c_void_p, C.c_size_t, C.POINTER(C.c_double)]
        L.oracle_apply_heritability.restype = C.c_double
        L.oracle_liability_order.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_liability_split.argtypes = [C.c_float, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_liability_labels.argtypes = [C.c_float, C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_liability_labels.restype = C.c_int
        L.oracle_scale_effects.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.oracle_scale_effects.restype = C.c_long
        L.oracle_synth_rows.argtypes = [C.c_uint64, C.c_int64, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
        L.oracle_synth_cols.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                        C.c_void_p, C.c_size_t]
        _LIB = L
    return _LIB


def ref() -> C.CDLL | None:
    """The reference's compiled libplinkio + harness, or None if not built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libsimu_ref.so")
        if not os.path.exists(path):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_dims.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        R.ref_read_all.argtypes = [C.c_char_p, C.c_void_p]
        R.ref_freq_pheno.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        R.ref_locus.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int), C.c_char_p, C.c_int,
                                C.POINTER(C.c_longlong), C.c_char_p, C.c_char_p, C.c_int]
        R.ref_sample.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_char_p, C.c_int]
        _REF = R
    return _REF


_REFSRC = None
SIMU_REF_BIN = os.path.join(_HERE, "_ref", "simu_ref")


def refsrc() -> C.CDLL | None:
    """The reference's UNMODIFIED src/source.cc compiled against oracle/boost_shim/
    (oracle/_ref/libsimu_refsrc.so): its literal find_freq / find_pheno / apply_* / main."""
    global _REFSRC
    if _REFSRC is None:
        path = os.path.join(_HERE, "_ref", "libsimu_refsrc.so")
        if not os.path.exists(path):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_src_last_error.restype = C.c_char_p
        R.ref_src_freq_pheno.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_src_apply_heritability.argtypes = [C.c_float, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p,
                                                 C.c_longlong, C.c_void_p]
        R.ref_src_apply_liability_threshold.argtypes = [C.c_float, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_longlong]
        _REFSRC = R
    return _REFSRC


def refsrc_freq_pheno(prefix: str, comp, x1, x2=None):
    """find_freq + find_pheno of the literal reference source on a bfile.
    -> (freq[M] with NaN for non-causal, pheno1, pheno2 or None, (t_freq, t_pheno) seconds)"""
    R = refsrc()
    comp = np.ascontiguousarray(comp, dtype=np.int32)
    m = comp.size
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    k2 = x2 is not None
    x2a = np.ascontiguousarray(x2, dtype=np.float64) if k2 else None
    n = sum(1 for _ in open(prefix + ".fam"))
    freq = np.zeros(m)
    p1 = np.zeros(n)
    p2 = np.zeros(n) if k2 else None
    secs = np.zeros(2)
    rc = R.ref_src_freq_pheno(os.fsencode(prefix), 2 if k2 else 1, _ptr(comp), m, _ptr(freq), _ptr(x1), _ptr(x2a),
                              _ptr(p1), _ptr(p2), _ptr(secs))
    if rc:
        raise RuntimeError((R.ref_src_last_error() or b"").decode())
    return freq, p1, p2, (float(secs[0]), float(secs[1]))


def refsrc_apply_heritability(hsq: float, seed: int, pheno, effect):
    """-> (pheno, effect, noise) after the literal apply_heritability with mt19937(seed)."""
    R = refsrc()
    p = np.array(pheno, dtype=np.float64)
    e = np.array(effect, dtype=np.float64)
    noise = np.zeros(p.size)
    if R.ref_src_apply_heritability(hsq, seed, _ptr(p), p.size, _ptr(e), e.size, _ptr(noise)):
        raise RuntimeError((R.ref_src_last_error() or b"").decode())
    return p, e, noise


def refsrc_apply_liability_threshold(k: float, ncas: int, ncon: int, seed: int, pheno):
    R = refsrc()
    p = np.array(pheno, dtype=np.float64)
    if R.ref_src_apply_liability_threshold(k, ncas, ncon, seed, _ptr(p), p.size):
        raise RuntimeError((R.ref_src_last_error() or b"").decode())
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- wrappers

def bed_header(b: bytes):
    v, o = C.c_int(), C.c_int()
    off = lib().oracle_bed_header(np.frombuffer(bytes(b[:3]).ljust(3, b"\0"), dtype=np.uint8).copy(),
                                  C.byref(v), C.byref(o))
    return v.value, o.value, off


def row_bytes(n: int) -> int:
    return int(lib().oracle_bed_row_bytes(n))


def unpack(packed: np.ndarray, n: int) -> np.ndarray:
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(n, dtype=np.uint8)
    lib().oracle_unpack_snps(_ptr(packed), _ptr(out), n)
    return out


def pack(codes: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    out = np.empty(row_bytes(codes.size), dtype=np.uint8)
    lib().oracle_pack_snps(_ptr(codes), _ptr(out), codes.size)
    return out


def find_freq(rows: np.ndarray, n: int, row_index=None):
    """rows: [R, stride] uint8.  Returns (counts[mc,4] int32, freq[mc] f64)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    idx = None if row_index is None else np.ascontiguousarray(row_index, dtype=np.int64)
    mc = rows.shape[0] if idx is None else idx.size
    counts = np.zeros((mc, 4), dtype=np.int32)
    freq = np.zeros(mc, dtype=np.float64)
    lib().oracle_find_freq(_ptr(rows), rows.shape[1], _ptr(idx), mc, n, _ptr(counts), _ptr(freq))
    return counts, freq


def find_pheno(rows: np.ndarray, n: int, freq, x1, x2=None, row_index=None, threads: int = 1, samples=None):
    """find_pheno of src/source.cc:951-1006.  threads > 1 splits the SAMPLE loop over worker threads
    (every sample's sum keeps the reference's ascending-SNP order, so the result is bit-identical to the
    single-threaded one); samples=(s0, s1) computes only that sample range (s0 % 4 == 0; the rest of the
    returned arrays stays 0)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    idx = None if row_index is None else np.ascontiguousarray(row_index, dtype=np.int64)
    mc = rows.shape[0] if idx is None else idx.size
    freq = np.ascontiguousarray(freq, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = None if x2 is None else np.ascontiguousarray(x2, dtype=np.float64)
    y1 = np.zeros(n, dtype=np.float64)
    y2 = None if x2 is None else np.zeros(n, dtype=np.float64)
    if threads <= 1 and samples is None:
        lib().oracle_find_pheno(_ptr(rows), rows.shape[1], _ptr(idx), mc, n, _ptr(freq), _ptr(x1), _ptr(x2),
                                _ptr(y1), _ptr(y2))
        return y1, y2
    s0, s1 = (0, n) if samples is None else samples
    assert s0 % 4 == 0 and 0 <= s0 <= s1 <= n
    L = lib()
    threads = max(1, min(threads, (s1 - s0 + 255) // 256))
    per = ((s1 - s0 + threads - 1) // threads + 3) // 4 * 4

    def work(t):
        a, b = s0 + t * per, min(s1, s0 + (t + 1) * per)
        if a < b:
            L.oracle_find_pheno_range(_ptr(rows), rows.shape[1], _ptr(idx), mc, a, b, _ptr(freq), _ptr(x1), _ptr(x2),
                                      _ptr(y1), _ptr(y2))

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return y1, y2


def apply_heritability(hsq: float, pheno: np.ndarray, noise: np.ndarray, effect: np.ndarray | None = None):
    """In place on copies; returns (pheno, effect, gen_coef, env_coef)."""
    pheno = np.array(pheno, dtype=np.float64)
    noise = np.ascontiguousarray(noise, dtype=np.float64)
    eff = None if effect is None else np.array(effect, dtype=np.float64)
    env = C.c_double()
    gen = lib().oracle_apply_heritability(hsq, _ptr(pheno), pheno.size, _ptr(noise), _ptr(eff),
                                          0 if eff is None else eff.size, C.byref(env))
    return pheno, eff, gen, env.value


def scale_effects(op: int, freq, effect1, effect2=None, component=None, s_pow1=None, s_pow2=None):
    """source.cc:1245-1262 (op 0), :1272-1286 (op 1), :1089-1094 (op 2) on copies.
    -> (effect1, effect2, bad) with bad = -1 or the first zero-het variant."""
    f = np.ascontiguousarray(freq, dtype=np.float64)
    x1 = np.array(effect1, dtype=np.float64)
    x2 = None if effect2 is None else np.array(effect2, dtype=np.float64)
    comp = None if component is None else np.ascontiguousarray(component, dtype=np.int32)
    sp1 = None if s_pow1 is None else np.ascontiguousarray(np.atleast_1d(s_pow1), dtype=np.float64)
    sp2 = None if s_pow2 is None else np.ascontiguousarray(np.atleast_1d(s_pow2), dtype=np.float64)
    bad = lib().oracle_scale_effects(op, _ptr(f), _ptr(x1), _ptr(x2), f.size, _ptr(comp), _ptr(sp1), _ptr(sp2))
    return x1, x2, int(bad)


def liability_order(pheno: np.ndarray) -> np.ndarray:
    pheno = np.ascontiguousarray(pheno, dtype=np.float64)
    idx = np.empty(pheno.size, dtype=np.int32)
    lib().oracle_liability_order(_ptr(pheno), pheno.size, _ptr(idx))
    return idx


def liability_split(k: float, n: int):
    a, b = C.c_int(), C.c_int()
    lib().oracle_liability_split(k, n, C.byref(a), C.byref(b))
    return a.value, b.value


def liability_labels(k, ncas, ncon, index, con_inds, cas_inds):
    n = index.size
    out = np.empty(n, dtype=np.float64)
    rc = lib().oracle_liability_labels(k, ncas, ncon, n,
                                       _ptr(np.ascontiguousarray(index, dtype=np.int32)),
                                       _ptr(np.ascontiguousarray(con_inds, dtype=np.int32)),
                                       _ptr(np.ascontiguousarray(cas_inds, dtype=np.int32)), _ptr(out))
    if rc:
        raise RuntimeError("ERROR: too many controls requested" if rc == -1 else "ERROR: too many cases requested")
    return out


def synth_rows(seed: int, first_row: int, n_rows: int, n: int, stride: int | None = None,
               threads: int = 0) -> np.ndarray:
    """Synthetic packed rows [n_rows, stride] (SURVEY.md 8d generator)."""
    stride = stride or row_bytes(n)
    out = np.zeros((n_rows, stride), dtype=np.uint8)
    threads = threads or min(os.cpu_count() or 1, 16)
    L = lib()
    if n_rows * n < 1 << 22 or threads == 1:
        L.oracle_synth_rows(seed, first_row, n_rows, n, _ptr(out), stride)
        return out
    chunk = (n_rows + threads - 1) // threads

    def work(t):
        lo, hi = t * chunk, min(n_rows, (t + 1) * chunk)
        if lo < hi:
            L.oracle_synth_rows(seed, first_row + lo, hi - lo, n, C.c_void_p(out.ctypes.data + lo * stride), stride)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def synth_cols(seed: int, s0: int, s1: int, first_row: int = 0, n_rows: int | None = None, row_ids=None,
               threads: int = 0) -> np.ndarray:
    """Sample columns [s0, s1) (s0 % 4 == 0) of synthetic rows, packed as a panel of s1 - s0 samples."""
    assert s0 % 4 == 0 and s1 > s0
    ids = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.int64)
    n_rows = ids.size if ids is not None else n_rows
    stride = row_bytes(s1 - s0)
    out = np.zeros((n_rows, stride), dtype=np.uint8)
    threads = threads or min(os.cpu_count() or 1, 16)
    L = lib()
    chunk = (n_rows + threads - 1) // threads

    def work(t):
        lo, hi = t * chunk, min(n_rows, (t + 1) * chunk)
        if lo < hi:
            L.oracle_synth_cols(seed, first_row + lo, None if ids is None else C.c_void_p(ids.ctypes.data + 8 * lo), hi - lo,
                                s0, s1, C.c_void_p(out.ctypes.data + lo * stride), stride)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def find_freq_threaded(rows: np.ndarray, n: int, threads: int = 0):
    """find_freq with the ROW loop split over worker threads (rows are independent: same bits)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    mc = rows.shape[0]
    counts = np.zeros((mc, 4), dtype=np.int32)
    freq = np.zeros(mc, dtype=np.float64)
    threads = max(1, min(threads or min(os.cpu_count() or 1, 16), mc))
    chunk = (mc + threads - 1) // threads
    L = lib()

    def work(t):
        lo, hi = t * chunk, min(mc, (t + 1) * chunk)
        if lo < hi:
            L.oracle_find_freq(C.c_void_p(rows.ctypes.data + lo * rows.shape[1]), rows.shape[1], None, hi - lo, n,
                               C.c_void_p(counts.ctypes.data + 16 * lo), C.c_void_p(freq.ctypes.data + 8 * lo))

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return counts, freq
