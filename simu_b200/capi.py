"""ctypes binding of the C ABI declared in ``include/simu_b200.h``.

This is the same stub a maintainer of the reference would write for a cgo /
ctypes / JNI binding: plain pointers and sizes, integer
status codes.  There is no Python or CPU fallback -- if the shared library is
missing, or no B200 is present, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# SIMU_B200_LIB selects another build of the SAME library (kernel-tuning A/B runs)
LIB_PATH = os.environ.get("SIMU_B200_LIB") or os.path.join(_PKG, "lib", "libsimu_b200.so")

OK, END, ERROR, EINVAL, ECUDA, ENOMEM, EFORMAT = range(7)
MODE_EXACT, MODE_FAST = 0, 1
LAYOUT_ROWS, LAYOUT_GROUP4, LAYOUT_SAMPLE16 = 0, 1, 2

# every symbol include/simu_b200.h declares: (name, restype, argtypes)
_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
SIGNATURES = {
    "simu_b200_version": (_i32, []),
    "simu_b200_device_count": (_i32, [C.POINTER(C.c_int)]),
    "simu_b200_open": (_i32, [C.POINTER(_vp), _i32, _i64]),
    "simu_b200_close": (None, [_vp]),
    "simu_b200_last_error": (C.c_char_p, [_vp]),
    "simu_b200_set_stream": (_i32, [_vp, _vp]),
    "simu_b200_sync": (_i32, [_vp]),
    "simu_b200_launch_count": (_i64, [_vp]),
    "simu_b200_row_bytes": (_i64, [_i64]),
    "simu_b200_row_stride": (_i64, [_vp]),
    "simu_b200_panel_rows": (_i64, [_vp]),
    "simu_b200_panel_device_ptr": (_vp, [_vp]),
    "simu_b200_panel_alloc": (_i32, [_vp, _i64]),
    "simu_b200_panel_alloc_layout": (_i32, [_vp, _i64, _i32]),
    "simu_b200_panel_layout": (_i32, [_vp]),
    "simu_b200_group_stride": (_i64, [_vp]),
    "simu_b200_panel_synth_rows": (_i32, [_vp, C.c_uint64, _vp, _i64, _i64]),
    "simu_b200_panel_free": (_i32, [_vp]),
    "simu_b200_panel_put_rows": (_i32, [_vp, _i64, _i64, _vp, _i64]),
    "simu_b200_panel_load_bed": (_i32, [_vp, C.c_char_p, _i64, _vp, _i64, _i64]),
    "simu_b200_panel_synth": (_i32, [_vp, C.c_uint64, _i64, _i64, _i64]),
    "simu_b200_panel_get_rows": (_i32, [_vp, _i64, _i64, _vp, _i64]),
    "simu_b200_find_freq": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "simu_b200_find_pheno": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32]),
    "simu_b200_find_pheno_accumulate": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32]),
    "simu_b200_mem_info": (_i32, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "simu_b200_panel_bytes": (_i64, [_vp, _i64, _i32]),
    "simu_b200_find_freq_pheno": (_i32, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                         C.POINTER(C.c_int64)]),
    "simu_b200_find_freq_pheno_dev": (_i32, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "simu_b200_apply_heritability": (_i32, [_vp, _f32, _vp, _i64, _vp, _vp, _i64,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "simu_b200_liability_order": (_i32, [_vp, _vp, _i64, _vp]),
    "simu_b200_liability_split": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "simu_b200_liability_split_dev": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "simu_b200_scale_effects": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, C.POINTER(C.c_int64)]),
    "simu_b200_scale_effects_dev": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "simu_b200_find_freq_dev": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "simu_b200_find_pheno_dev": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32]),
    "simu_b200_sumsq_dev": (_i32, [_vp, _vp, _i64, _vp]),
    "simu_b200_heritability_dev": (_i32, [_vp, _f32, _vp, _i64, _vp, _vp, _vp]),
    "simu_b200_apply_heritability_dev": (_i32, [_vp, _f32, _vp, _i64, _vp, _vp]),
    "simu_b200_liability_order_dev": (_i32, [_vp, _vp, _i64, _vp]),
    "simu_b200_peer_init": (_i32, [_vp, _i32, _i32, _i64, _vp]),
    "simu_b200_peer_connect": (_i32, [_vp, _vp]),
    "simu_b200_peer_allreduce_dev": (_i32, [_vp, _vp, _i64]),
    "simu_b200_peer_close": (None, [_vp]),
    "simu_b200_missing_index": (_i32, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "simu_b200_last_kernel_ms": (_i32, [_vp, _i32, C.POINTER(C.c_float)]),
    "simu_b200_profile_enable": (_i32, [_vp, _i32]),
    "simu_b200_profile_read": (_i32, [_vp, _i32, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
}

_LIB = None


class SimuB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def load() -> C.CDLL:
    """Load libsimu_b200.so and bind every declared symbol; raise if absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m simu_b200.build` "
                "(or __graft_entry__.build()).  simu_b200 has no CPU or Python fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One run on one GPU: owns the HBM panel, the streams and the scratch."""

    def __init__(self, num_samples: int, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.simu_b200_open(C.byref(h), device, num_samples)
        if rc != OK:
            raise SimuB200Error(rc, (self.lib.simu_b200_last_error(None) or b"").decode())
        self.h = h
        self.num_samples = int(num_samples)
        self.device = device
        self.row_bytes = int(self.lib.simu_b200_row_bytes(num_samples))
        self.row_stride = int(self.lib.simu_b200_row_stride(h))

    # -- plumbing -----------------------------------------------------------
    def _check(self, rc: int):
        if rc != OK:
            raise SimuB200Error(rc, (self.lib.simu_b200_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.simu_b200_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.simu_b200_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self._check(self.lib.simu_b200_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.simu_b200_launch_count(self.h))

    def missing_index(self):
        """(state, entries): 0 not built, 1 sparse list of `entries` missing genotypes, 2 dense plane (simu_b200.h)."""
        st, n = C.c_int(0), C.c_int64(0)
        self._check(self.lib.simu_b200_missing_index(self.h, C.byref(st), C.byref(n)))
        return st.value, n.value

    def last_kernel_ms(self, which: int) -> float:
        ms = C.c_float()
        self._check(self.lib.simu_b200_last_kernel_ms(self.h, which, C.byref(ms)))
        return ms.value

    def profile_enable(self, on: bool = True):
        self._check(self.lib.simu_b200_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, kernel_id: int):
        """-> (total_ms, launches) of the recorded launches of one kernel."""
        ms, n = C.c_float(), C.c_int()
        self._check(self.lib.simu_b200_profile_read(self.h, kernel_id, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- panel --------------------------------------------------------------
    @property
    def panel_rows(self) -> int:
        return int(self.lib.simu_b200_panel_rows(self.h))

    @property
    def panel_ptr(self) -> int:
        return int(self.lib.simu_b200_panel_device_ptr(self.h) or 0)

    def panel_alloc(self, rows: int, layout: int = LAYOUT_ROWS):
        """layout: LAYOUT_ROWS (row lists allowed), or a compact causal panel (rows=None): LAYOUT_GROUP4
        (lookup-table FAST kernel) or LAYOUT_SAMPLE16 (tensor-core FAST kernel)."""
        self._check(self.lib.simu_b200_panel_alloc_layout(self.h, rows, layout))

    @property
    def layout(self) -> int:
        return int(self.lib.simu_b200_panel_layout(self.h))

    @property
    def group_stride(self) -> int:
        return int(self.lib.simu_b200_group_stride(self.h))

    def panel_free(self):
        self._check(self.lib.simu_b200_panel_free(self.h))

    def panel_put_rows(self, dst_row: int, host_rows: np.ndarray):
        """host_rows: uint8 [n_rows, stride>=row_bytes], C-contiguous."""
        a = np.ascontiguousarray(host_rows, dtype=np.uint8)
        if a.ndim != 2:
            raise ValueError("host_rows must be 2-D [rows, bytes]")
        self._check(self.lib.simu_b200_panel_put_rows(self.h, dst_row, a.shape[0], _np_ptr(a), a.shape[1]))

    def panel_put_rows_ptr(self, dst_row: int, n_rows: int, host_ptr: int, host_stride: int):
        self._check(self.lib.simu_b200_panel_put_rows(self.h, dst_row, n_rows, C.c_void_p(host_ptr), host_stride))

    def panel_get_rows(self, src_row: int, n_rows: int) -> np.ndarray:
        out = np.empty((n_rows, self.row_bytes), dtype=np.uint8)
        self._check(self.lib.simu_b200_panel_get_rows(self.h, src_row, n_rows, _np_ptr(out), self.row_bytes))
        return out

    def panel_get_rows_ptr(self, src_row: int, n_rows: int, host_ptr: int, host_stride: int):
        self._check(self.lib.simu_b200_panel_get_rows(self.h, src_row, n_rows, C.c_void_p(host_ptr), host_stride))

    def panel_load_bed(self, bed_path: str, num_loci: int, file_rows=None, dst_row: int = 0, n_rows: int | None = None):
        fr = None if file_rows is None else np.ascontiguousarray(file_rows, dtype=np.int64)
        n = (num_loci if n_rows is None else n_rows) if fr is None else fr.size
        self._check(self.lib.simu_b200_panel_load_bed(self.h, os.fsencode(bed_path), num_loci, _np_ptr(fr), n, dst_row))

    def panel_synth(self, seed: int, first_row: int, n_rows: int, dst_row: int = 0):
        self._check(self.lib.simu_b200_panel_synth(self.h, seed, first_row, n_rows, dst_row))

    def panel_synth_rows(self, seed: int, row_ids, dst_row: int = 0):
        ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        self._check(self.lib.simu_b200_panel_synth_rows(self.h, seed, _np_ptr(ids), ids.size, dst_row))

    # -- hot path, host buffers ----------------------------------------------
    def find_freq(self, rows=None, mc: int | None = None):
        """-> (counts [mc,4] int32, freq [mc] float64).  source.cc:843-885."""
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        mc = (self.panel_rows if mc is None else mc) if r is None else r.size
        counts = np.zeros((mc, 4), dtype=np.int32)
        freq = np.zeros(mc, dtype=np.float64)
        self._check(self.lib.simu_b200_find_freq(self.h, _np_ptr(r), mc, _np_ptr(counts), _np_ptr(freq)))
        return counts, freq

    def find_pheno(self, freq, effect1, effect2=None, rows=None, mode: int = MODE_EXACT):
        """-> (pheno1, pheno2 or None).  source.cc:951-1006."""
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        f = np.ascontiguousarray(freq, dtype=np.float64)
        x1 = np.ascontiguousarray(effect1, dtype=np.float64)
        x2 = None if effect2 is None else np.ascontiguousarray(effect2, dtype=np.float64)
        mc = f.size
        if x1.size != mc or (x2 is not None and x2.size != mc) or (r is not None and r.size != mc):
            raise ValueError("freq, effects and rows must have one entry per causal SNP")
        y1 = np.zeros(self.num_samples, dtype=np.float64)
        y2 = None if x2 is None else np.zeros(self.num_samples, dtype=np.float64)
        self._check(self.lib.simu_b200_find_pheno(self.h, _np_ptr(r), mc, _np_ptr(f), _np_ptr(x1), _np_ptr(x2),
                                                  _np_ptr(y1), _np_ptr(y2), mode))
        return y1, y2

    def find_pheno_accumulate(self, freq, effect1, effect2, pheno1, pheno2=None, mode: int = MODE_EXACT):
        """pheno += the contributions of the current (SAMPLE16) panel's rows, in place; see include/simu_b200.h."""
        f = np.ascontiguousarray(freq, dtype=np.float64)
        x1 = np.ascontiguousarray(effect1, dtype=np.float64)
        x2 = None if effect2 is None else np.ascontiguousarray(effect2, dtype=np.float64)
        if pheno1.dtype != np.float64 or not pheno1.flags.c_contiguous or (pheno2 is not None and pheno2.dtype != np.float64):
            raise ValueError("pheno arrays must be contiguous float64 (they are updated in place)")
        self._check(self.lib.simu_b200_find_pheno_accumulate(self.h, f.size, _np_ptr(f), _np_ptr(x1), _np_ptr(x2),
                                                             _np_ptr(pheno1), _np_ptr(pheno2), mode))

    def mem_info(self):
        """-> (free_bytes, total_bytes) of the device; the current panel counts as free."""
        f, t = C.c_int64(), C.c_int64()
        self._check(self.lib.simu_b200_mem_info(self.h, C.byref(f), C.byref(t)))
        return f.value, t.value

    def panel_bytes(self, rows: int, layout: int) -> int:
        return int(self.lib.simu_b200_panel_bytes(self.h, rows, layout))

    def find_freq_pheno(self, raw1, raw2=None, scale_op: int = -1, component=None, s_pow1=None, s_pow2=None, mc=None):
        """find_freq + effect scaling + find_pheno (FAST) in one pass (GROUP4 panels).
        -> (counts, freq, effect1, effect2 or None, pheno1, pheno2 or None)."""
        r1 = np.ascontiguousarray(raw1, dtype=np.float64)
        r2 = None if raw2 is None else np.ascontiguousarray(raw2, dtype=np.float64)
        mc = r1.size if mc is None else mc
        comp = None if component is None else np.ascontiguousarray(component, dtype=np.int32)
        sp1 = None if s_pow1 is None else np.ascontiguousarray(np.atleast_1d(s_pow1), dtype=np.float64)
        sp2 = None if s_pow2 is None else np.ascontiguousarray(np.atleast_1d(s_pow2), dtype=np.float64)
        ncomp = max(1, 0 if sp1 is None else sp1.size)
        counts = np.zeros((mc, 4), dtype=np.int32)
        freq = np.zeros(mc, dtype=np.float64)
        x1 = np.zeros(mc, dtype=np.float64)
        x2 = None if r2 is None else np.zeros(mc, dtype=np.float64)
        y1 = np.zeros(self.num_samples, dtype=np.float64)
        y2 = None if r2 is None else np.zeros(self.num_samples, dtype=np.float64)
        bad = C.c_int64(-1)
        rc = self.lib.simu_b200_find_freq_pheno(self.h, mc, _np_ptr(r1), _np_ptr(r2), scale_op, _np_ptr(comp), _np_ptr(sp1),
                                                _np_ptr(sp2), ncomp, _np_ptr(counts), _np_ptr(freq), _np_ptr(x1), _np_ptr(x2),
                                                _np_ptr(y1), _np_ptr(y2), C.byref(bad))
        self.last_bad_index = bad.value
        self._check(rc)
        return counts, freq, x1, x2, y1, y2

    def apply_heritability(self, hsq: float, pheno, noise, effect=None):
        """-> (pheno, effect, gen_coef, env_coef).  source.cc:1008-1029."""
        y = np.array(pheno, dtype=np.float64)
        e = np.ascontiguousarray(noise, dtype=np.float64)
        if e.size != y.size:
            raise ValueError("noise must have one draw per sample")
        x = None if effect is None else np.array(effect, dtype=np.float64)
        g, v = C.c_double(), C.c_double()
        self._check(self.lib.simu_b200_apply_heritability(self.h, hsq, _np_ptr(y), y.size, _np_ptr(e), _np_ptr(x),
                                                          0 if x is None else x.size, C.byref(g), C.byref(v)))
        return y, x, g.value, v.value

    def scale_effects(self, op: int, freq, effect1, effect2=None, component=None, s_pow1=None, s_pow2=None):
        """op 0: *= sqrt(pow(het, S)); 1: /= sqrt(het); 2: *= sqrt(het).  source.cc:1245-1262, :1272-1286, :1089-1094.
        -> (effect1, effect2 or None); raises SimuB200Error(ERROR) naming the first zero-het variant."""
        f = np.ascontiguousarray(freq, dtype=np.float64)
        x1 = np.array(effect1, dtype=np.float64)
        x2 = None if effect2 is None else np.array(effect2, dtype=np.float64)
        comp = None if component is None else np.ascontiguousarray(component, dtype=np.int32)
        sp1 = None if s_pow1 is None else np.ascontiguousarray(np.atleast_1d(s_pow1), dtype=np.float64)
        sp2 = None if s_pow2 is None else np.ascontiguousarray(np.atleast_1d(s_pow2), dtype=np.float64)
        ncomp = max(1, 0 if sp1 is None else sp1.size)
        bad = C.c_int64(-1)
        rc = self.lib.simu_b200_scale_effects(self.h, op, _np_ptr(f), _np_ptr(x1), _np_ptr(x2), f.size, _np_ptr(comp),
                                              _np_ptr(sp1), _np_ptr(sp2), ncomp, C.byref(bad))
        self.last_bad_index = bad.value
        self._check(rc)
        return x1, x2

    def liability_order(self, pheno) -> np.ndarray:
        y = np.ascontiguousarray(pheno, dtype=np.float64)
        idx = np.empty(y.size, dtype=np.int32)
        self._check(self.lib.simu_b200_liability_order(self.h, _np_ptr(y), y.size, _np_ptr(idx)))
        return idx

    def liability_split(self, pheno, max_ncon: int) -> np.ndarray:
        """-> uint8 flags: 0 = control pool (the max_ncon lowest liabilities), 1 = case pool."""
        y = np.ascontiguousarray(pheno, dtype=np.float64)
        flags = np.empty(y.size, dtype=np.uint8)
        self._check(self.lib.simu_b200_liability_split(self.h, _np_ptr(y), y.size, max_ncon, _np_ptr(flags)))
        return flags

    # -- hot path, device pointers (ints) -------------------------------------
    def find_freq_dev(self, d_rows, mc, d_counts, d_freq):
        self._check(self.lib.simu_b200_find_freq_dev(self.h, C.c_void_p(d_rows or 0), mc, C.c_void_p(d_counts or 0),
                                                     C.c_void_p(d_freq or 0)))

    def find_pheno_dev(self, d_rows, mc, d_freq, d_x1, d_x2, d_y1, d_y2, mode=MODE_EXACT):
        self._check(self.lib.simu_b200_find_pheno_dev(self.h, C.c_void_p(d_rows or 0), mc, C.c_void_p(d_freq),
                                                      C.c_void_p(d_x1), C.c_void_p(d_x2 or 0), C.c_void_p(d_y1),
                                                      C.c_void_p(d_y2 or 0), mode))

    def find_freq_pheno_dev(self, mc, d_raw1, d_raw2, scale_op, d_comp, d_spow, ncomp, d_counts, d_freq, d_x1, d_x2,
                            d_y1, d_y2, d_bad):
        v = lambda p: C.c_void_p(p or 0)
        self._check(self.lib.simu_b200_find_freq_pheno_dev(self.h, mc, v(d_raw1), v(d_raw2), scale_op, v(d_comp), v(d_spow),
                                                           ncomp, v(d_counts), v(d_freq), v(d_x1), v(d_x2), v(d_y1),
                                                           v(d_y2), v(d_bad)))

    # -- multi-GPU exchange over NVLink peer memory --------------------------------
    def peer_init(self, rank: int, world: int, count: int) -> bytes:
        """-> the 64-byte IPC handle of this rank's symmetric buffer (exchange it, then peer_connect)."""
        h = C.create_string_buffer(64)
        self._check(self.lib.simu_b200_peer_init(self.h, rank, world, count, h))
        return h.raw

    def peer_connect(self, handles: list[bytes]):
        blob = b"".join(handles)
        self._check(self.lib.simu_b200_peer_connect(self.h, C.c_char_p(blob)))

    def peer_allreduce_dev(self, d_y, count: int):
        self._check(self.lib.simu_b200_peer_allreduce_dev(self.h, C.c_void_p(d_y), count))

    def sumsq_dev(self, d_y, n, d_out):
        self._check(self.lib.simu_b200_sumsq_dev(self.h, C.c_void_p(d_y), n, C.c_void_p(d_out)))

    def apply_heritability_dev(self, hsq, d_y, n, d_noise, d_coef=None):
        self._check(self.lib.simu_b200_apply_heritability_dev(self.h, hsq, C.c_void_p(d_y), n, C.c_void_p(d_noise),
                                                              C.c_void_p(d_coef or 0)))

    def heritability_dev(self, hsq, d_y, n, d_noise, d_sumsq, d_coef):
        self._check(self.lib.simu_b200_heritability_dev(self.h, hsq, C.c_void_p(d_y), n, C.c_void_p(d_noise),
                                                        C.c_void_p(d_sumsq), C.c_void_p(d_coef or 0)))

    def liability_split_dev(self, d_pheno, n, max_ncon, d_is_case):
        self._check(self.lib.simu_b200_liability_split_dev(self.h, C.c_void_p(d_pheno), n, max_ncon, C.c_void_p(d_is_case)))

    def liability_order_dev(self, d_pheno, n, d_index):
        self._check(self.lib.simu_b200_liability_order_dev(self.h, C.c_void_p(d_pheno), n, C.c_void_p(d_index)))
