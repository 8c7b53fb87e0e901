"""Host-side mirror of the reference's hot-path functions, over the C ABI.

Same names, argument meaning and error behaviour as
/root/reference/src/source.cc:

    find_freq                  :843-885
    find_pheno                 :951-1006
    apply_heritability         :1008-1029
    apply_liability_threshold  :1031-1055

`component_per_variant` is the reference's vector with -1 for non-causal
variants (:759-769); frequency / effect vectors have one entry per variant.
Everything heavy happens in libsimu_b200.so on the GPU; this module only
gathers the causal subset, shards it over ranks and scatters results back.

Multi-GPU: causal SNPs are split into contiguous slices
balanced by causal-row count; each rank stages and processes only its slice;
allele counts are exchanged with one all-gather of the (tiny) freq slices and
the partial N x K phenotypes are combined with ONE all-reduce(sum, float64).
With torch.distributed uninitialised (or world size 1) no collective runs.
"""
from __future__ import annotations

import math

import numpy as np

from . import capi
from .plink import PlinkFiles

FLT_EPSILON = float(np.finfo(np.float32).eps)


# ------------------------------------------------------------------ sharding

def shard_bounds(mc: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced slices of the ascending causal list."""
    base, rem = divmod(mc, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class _Dist:
    """torch.distributed facade; a no-op when not initialised."""

    def __init__(self):
        self.rank, self.world, self.dist = 0, 1, None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.rank, self.world, self.dist = dist.get_rank(), dist.get_world_size(), dist
        except Exception:
            pass

    def all_reduce_sum(self, arr: np.ndarray) -> np.ndarray:
        if self.dist is None or self.world == 1:
            return arr
        import torch
        backend = self.dist.get_backend()
        t = torch.from_numpy(arr)
        if backend == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def all_gather_concat(self, arr: np.ndarray, sizes: list[int]) -> np.ndarray:
        """Concatenate per-rank slices (rank order) of known sizes."""
        if self.dist is None or self.world == 1:
            return arr
        import torch
        backend = self.dist.get_backend()
        dev = "cuda" if backend == "nccl" else "cpu"
        mx = max(sizes) if sizes else 0
        pad = np.zeros((mx,) + arr.shape[1:], dtype=arr.dtype)
        pad[: arr.shape[0]] = arr
        t = torch.from_numpy(pad).to(dev)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return np.concatenate([o.cpu().numpy()[: sizes[r]] for r, o in enumerate(outs)], axis=0)


# --------------------------------------------------------------- the session

class HotPath:
    """State the reference keeps implicitly in its PioFiles cursor: which
    causal rows are staged in HBM and in what order."""

    def __init__(self, ctx: capi.Context, pio_files: PlinkFiles | None = None):
        self.ctx = ctx
        self.pio_files = pio_files
        self.dist = _Dist()
        self._staged_key = None
        self._causal = None           # ascending global variant indices (all ranks)
        self._lo = self._hi = 0       # this rank's slice of the causal list

    def _causal_indices(self, simu_options, component_per_variant) -> np.ndarray:
        comp = np.asarray(component_per_variant)
        nv = int(simu_options["num_variants"] if isinstance(simu_options, dict) else simu_options.num_variants)
        if comp.size < nv:   # stream shorter than announced (source.cc:855-858)
            raise RuntimeError(f"ERROR: expect to find {nv} variants across input files; found only {comp.size}. "
                               "Are input Plink files corrupted?")
        if comp.size > nv:   # source.cc:879-882
            raise RuntimeError(f"ERROR: expect to find {nv} variants across input files; more variants found. "
                               "Are input Plink files corrupted?")
        return np.flatnonzero(comp != -1).astype(np.int64)

    def _stage(self, causal: np.ndarray):
        """Stage this rank's slice of the causal rows (once per causal set)."""
        key = (causal.size, int(causal[:1].sum()), int(causal[-1:].sum()), hash(causal.tobytes()))
        if key == self._staged_key:
            return
        self._lo, self._hi = shard_bounds(causal.size, self.dist.world)[self.dist.rank]
        if self.pio_files is not None:
            if self.pio_files.num_samples != self.ctx.num_samples:
                raise RuntimeError("ERROR: Inconsistent number of subjects in --bfile-chr input")
            self.pio_files.stage_causal_rows(self.ctx, causal[self._lo:self._hi])
        self._causal, self._staged_key = causal, key

    # -- find_freq ------------------------------------------------------------
    def find_freq(self, simu_options, component_per_variant, return_counts: bool = False):
        """-> freq_vec (NaN for non-causal variants, source.cc:850)."""
        causal = self._causal_indices(simu_options, component_per_variant)
        nv = np.asarray(component_per_variant).size
        freq_vec = np.full(nv, np.nan, dtype=np.float64)
        counts_vec = np.zeros((nv, 4), dtype=np.int32)
        if causal.size:
            self._stage(causal)
            m_local = self._hi - self._lo
            counts, freq = self.ctx.find_freq(mc=m_local) if m_local else (np.zeros((0, 4), np.int32), np.zeros(0))
            sizes = [hi - lo for lo, hi in shard_bounds(causal.size, self.dist.world)]
            freq = self.dist.all_gather_concat(freq, sizes)
            counts = self.dist.all_gather_concat(counts, sizes)
            freq_vec[causal] = freq
            counts_vec[causal] = counts
        return (freq_vec, counts_vec) if return_counts else freq_vec

    # -- find_pheno -----------------------------------------------------------
    def find_pheno(self, simu_options, component_per_variant, freq_vec, effect1_per_variant,
                   effect2_per_variant=None, mode: int = capi.MODE_EXACT):
        """-> (pheno1_per_sample, pheno2_per_sample or None)."""
        num_traits = int(simu_options["num_traits"] if isinstance(simu_options, dict) else simu_options.num_traits)
        bivariate = num_traits == 2                                      # source.cc:959
        causal = self._causal_indices(simu_options, component_per_variant)
        n = self.ctx.num_samples
        y1 = np.zeros(n, dtype=np.float64)
        y2 = np.zeros(n, dtype=np.float64) if bivariate else None
        if causal.size == 0:
            return y1, y2
        self._stage(causal)
        mine = causal[self._lo:self._hi]
        if mine.size:
            f = np.asarray(freq_vec, dtype=np.float64)[mine]
            x1 = np.asarray(effect1_per_variant, dtype=np.float64)[mine]
            x2 = np.asarray(effect2_per_variant, dtype=np.float64)[mine] if bivariate else None
            y1, y2 = self.ctx.find_pheno(f, x1, x2, rows=None, mode=mode)
        if self.dist.world > 1:
            stacked = np.stack([y1, y2]) if bivariate else y1[None]
            stacked = self.dist.all_reduce_sum(np.ascontiguousarray(stacked))
            y1, y2 = stacked[0], (stacked[1] if bivariate else None)
        return y1, y2

    # -- apply_heritability ----------------------------------------------------
    def apply_heritability(self, hsq: float, noise, pheno_per_sample, effect_per_variant):
        """`noise` = the N normal draws the caller's RNG made (source.cc:1014-1016).
        -> (pheno, effect, gen_coef, env_coef)"""
        return self.ctx.apply_heritability(float(np.float32(hsq)), pheno_per_sample, noise, effect_per_variant)

    # -- apply_liability_threshold ---------------------------------------------
    def apply_liability_threshold(self, k: float, ncas: int, ncon: int, shuffle, pheno_per_sample):
        """`shuffle(list)` is the caller's RNG-driven std::random_shuffle, called
        for the control pool then the case pool (source.cc:1046-1047).
        -> labels (1 control, 2 case, -9 unused) as float64."""
        y = np.asarray(pheno_per_sample, dtype=np.float64)
        n = int(y.size)
        max_ncas = int(math.floor(float(np.float32(k) * np.float32(n))))  # :1038 (float product)
        max_ncon = n - max_ncas                                           # :1039
        con_inds = list(range(0, max_ncon))                               # :1042
        cas_inds = list(range(max_ncon, n))                               # :1043
        con_inds = shuffle(con_inds)
        cas_inds = shuffle(cas_inds)
        if len(con_inds) < ncon:
            raise RuntimeError("ERROR: too many controls requested")      # :1049
        if len(cas_inds) < ncas:
            raise RuntimeError("ERROR: too many cases requested")         # :1050
        if ncon == max_ncon and ncas == max_ncas:
            # every sample of both pools is labelled (the default): the shuffles only consumed RNG
            # state, and the pools follow from the order statistic alone -- radix select, no sort
            return np.where(self.ctx.liability_split(y, max_ncon) != 0, 2.0, 1.0)
        index = self.ctx.liability_order(y)                               # :1032-1034
        out = np.full(n, -9.0)                                            # :1052
        out[index[np.asarray(con_inds[:ncon], dtype=np.int64)]] = 1       # :1053
        out[index[np.asarray(cas_inds[:ncas], dtype=np.int64)]] = 2       # :1054
        return out
