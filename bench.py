#!/usr/bin/env python
"""bench.py -- phenotype-synthesis hot path (y = Gx + e over packed PLINK .bed).

One "step" = one pass of the hot path over one batch of synthetic input:
    find_freq (allele counts) -> find_pheno (FAST LUT kernel) ->
    [N>1: one NCCL all-reduce of the partial N x K phenotypes] ->
    apply_heritability (sum of squares + scale-and-add-noise) per trait
    [-> liability sort for --cc workloads].

Metric (BASELINE.json): genotype x effect updates per second = N * Mc * K per
step, whole job over all GPUs; packed-.bed GB/s and the HBM roofline fraction of
the dominant kernel ride along.  Default workload = BASELINE configs[1]:
100K individuals x 1M SNPs, 100K causal SNPs, 2 traits, --gcta-sigma effects; the
causal rows (2.5 GB, all the reference ever reads) are resident in HBM as a compact
GROUP4 panel, as the drop-in CLI stages them (--layout rows: the whole 25 GB slice
row-major + a causal row list).  Data is synthetic (SURVEY.md 8d generator, panel seed
20240901) and generated on the device.

`--impl reference` times the reference's CPU algorithm (the oracle port of
src/source.cc:843-1029, single-threaded like the reference) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PANEL_SEED = 20240901

WORKLOADS = {
    # name: N, M rows resident per GPU, Mc causal per GPU, K traits, cc, description
    "config1": dict(n=10_000, m=100_000, mc=1_000, k=1, cc=False,
                    desc="--qt 10K x 100K, --causal-n 1000 --hsq 0.5 --seed 1"),
    "config2": dict(n=100_000, m=1_000_000, mc=100_000, k=2, cc=False,
                    desc="--qt --num-traits 2 --rg 0.5 --gcta-sigma --norm-effect, 100K x 1M, --causal-n 100000"),
    "config3": dict(n=100_000, m=1_000_000, mc=10_000, k=1, cc=True,
                    desc="--cc --k 0.1, 100K x 1M, --causal-n 10000"),
    "config4": dict(n=100_000, m=1_375_000, mc=1_375_000, k=1, cc=False,
                    desc="--bfile-chr @, 100K x 11M fully polygenic, 1/8 of the SNPs per GPU (34.5 GB)"),
    "config5": dict(n=500_000, m=1_000_000, mc=100_000, k=2, cc=False,
                    desc="--causal-variants, 2 traits, 500K x 1M (125 GB), 100K causal"),
}
HSQ = (0.5, 0.7)
RG = 0.5


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def _causal_rows(w, rank):
    rng = np.random.default_rng(1 + rank)            # --seed 1 (+rank: each GPU owns its own slice)
    if w["mc"] == w["m"]:
        return np.arange(w["m"], dtype=np.int32)
    return np.sort(rng.choice(w["m"], w["mc"], replace=False)).astype(np.int32)


def _raw_effects(w, rank):
    """Bivariate normal effect draws (source.cc:940-948), before the --gcta-sigma scaling."""
    rng = np.random.default_rng(1000 + rank)
    z1, z2 = rng.standard_normal(w["mc"]), rng.standard_normal(w["mc"])
    return z1, ((RG * z1 + np.sqrt(1.0 - RG * RG) * z2) if w["k"] == 2 else None)


def _effects(w, freq, rank):
    """... with --gcta-sigma scaling (:1257-1262)."""
    r1, r2 = _raw_effects(w, rank)
    het = np.abs(2.0 * freq * (1.0 - freq))
    scale = np.sqrt(np.power(np.maximum(het, 1e-12), -1.0))
    return r1 * scale, (r2 * scale if w["k"] == 2 else None)


def alg_bytes(w):
    B = (w["n"] + 3) // 4
    counts = w["mc"] * (B + 16 + 8)
    pheno = w["mc"] * (B + 12 + 8 * w["k"]) + 8 * w["n"] * w["k"]
    return B, counts, pheno


# ------------------------------------------------------------------ clocks

class ClockSampler:
    """SM clock, power and throttle reasons of one GPU, sampled IN-PROCESS through NVML every 5 ms by a thread
    (an `nvidia-smi -lms` child process needs ~100 ms for its first row, longer than a short timed region)."""

    _BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int, period: float = 0.005):
        self.samples, self.period, self._stop, self.t = [], period, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._one()                                   # at least one sample exists before any timed region starts
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:                            # no NVML: the line says so instead of inventing clocks
            self.nv, self.err = None, repr(e)

    def _one(self):
        nv = self.nv
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        try:
            power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:
            power = None
        self.samples.append((time.time(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(reasons), power))

    def _run(self):
        while not self._stop.wait(self.period):
            try:
                self._one()
            except Exception:
                break

    def stop(self):
        self._stop.set()

    def summary(self, t0, t1):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": getattr(self, "err", "no NVML")}
        rows = [s for s in self.samples if t0 <= s[0] <= t1]
        inside = len(rows)
        if not rows:                                      # region shorter than the period: the nearest samples around it
            rows = sorted(self.samples, key=lambda s: abs(s[0] - 0.5 * (t0 + t1)))[:3]
        reasons = sorted(name for name, bit in self._BAD.items() if any(r[2] & bit for r in rows))
        power = [r[3] for r in rows if r[3] is not None]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(rows), "samples_inside_timed_region": inside,
                "power_w_max": max(power) if power else None, "source": "NVML in-process, 5 ms period"}


# ------------------------------------------------------------ reference arm

def cpu_reference_time(rows_host, n, k, hsq=HSQ):
    """One pass of the reference's CPU hot path over `rows_host` ([s, ceil(n/4)] packed rows).

    kind "reference": the literal find_freq / find_pheno / apply_heritability of the UNMODIFIED
    /root/reference/src/source.cc (compiled against oracle/boost_shim into
    oracle/_ref/libsimu_refsrc.so), reading the rows through the reference's own libplinkio from
    a temporary bfile, single-threaded as upstream.  Falls back to the line-by-line C port
    (kind "port") only if that library did not travel.  -> (seconds, kind)"""
    import tempfile
    from oracle import oracle
    mc = rows_host.shape[0]
    rng = np.random.default_rng(7)
    x1 = rng.standard_normal(mc)
    x2 = rng.standard_normal(mc) if k == 2 else None
    if oracle.refsrc() is not None:
        from simu_b200 import plink
        with tempfile.TemporaryDirectory(prefix="simu_ref_") as d:
            prefix = os.path.join(d, "sample")
            plink.write_bed(prefix + ".bed", rows_host, n)
            plink.write_bim(prefix + ".bim", mc)
            plink.write_fam(prefix + ".fam", n)
            comp = np.zeros(mc, dtype=np.int32)
            _, y1, y2, (t_freq, t_pheno) = oracle.refsrc_freq_pheno(prefix, comp, x1, x2)    # source.cc:843-885, :951-1006
        t0 = time.perf_counter()
        oracle.refsrc_apply_heritability(hsq[0], 1, y1, x1)                                   # source.cc:1008-1029
        if k == 2:
            oracle.refsrc_apply_heritability(hsq[1], 2, y2, x2)
        return t_freq + t_pheno + (time.perf_counter() - t0), "reference"
    t0 = time.perf_counter()
    _, freq = oracle.find_freq(rows_host, n)
    y1, y2 = oracle.find_pheno(rows_host, n, freq, x1, x2)
    noise = rng.standard_normal(n)
    oracle.apply_heritability(hsq[0], y1, noise, x1)
    if k == 2:
        oracle.apply_heritability(hsq[1], y2, noise, x2)
    return time.perf_counter() - t0, "port"


def run_reference(args, w):
    """The reference's CPU implementation of the path, single-threaded as upstream
    (/root/reference has no threads: SURVEY.md 0.2): the literal functions of the unmodified
    src/source.cc from oracle/_ref (see cpu_reference_time)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    n, k = w["n"], w["k"]
    causal = _causal_rows(w, 0)
    probe_rows = 64
    rows = np.concatenate([oracle.synth_rows(PANEL_SEED, int(r), 1, n) for r in causal[:probe_rows]])
    t_probe, kind = cpu_reference_time(rows, n, k)
    budget = 100.0
    per_row = max(t_probe / probe_rows, 1e-7)
    s = int(max(probe_rows, min(w["mc"], budget / ((args.steps + args.warmup) * per_row))))
    # synthesize the sample: the first s causal rows of the same panel the GPU arm uses
    rows = np.empty((s, rows.shape[1]), np.uint8)
    runs = np.split(np.arange(s), np.flatnonzero(np.diff(causal[:s]) != 1) + 1) if s > 1 else [np.arange(s)]
    for run in runs:
        rows[run[0]:run[-1] + 1] = oracle.synth_rows(PANEL_SEED, int(causal[run[0]]), run.size, n)
    for _ in range(args.warmup):
        cpu_reference_time(rows, n, k)
    dt = sum(cpu_reference_time(rows, n, k)[0] for _ in range(args.steps)) / args.steps
    value = n * s * k / dt
    sample = f"first {s} of {w['mc']} causal SNP rows x {n} samples per step (same panel, same row list)"
    line = {
        "impl": "reference", "metric": "genotype_effect_updates_per_s", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["desc"], "n_samples": n, "n_causal": w["mc"],
                   "traits": k, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "updates/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores": os.cpu_count(),
                         "what": "unmodified src/source.cc find_freq+find_pheno+apply_heritability via its own libplinkio "
                                 "(oracle/_ref, Boost stand-in headers)" if kind == "reference" else "line-by-line C port"},
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "bed_gbs": s * ((n + 3) // 4) * 2 / dt / 1e9,
    }
    emit(line)


# ------------------------------------------------------------------ our arm

class _DevMem:
    """Zero-copy torch view of library-owned device memory (plumbing only)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _roofline(kernel, alg_bytes_per_launch, ms, n_launches, peaks, peak_src, traffic):
    if not n_launches:
        return None
    ach = alg_bytes_per_launch / (ms / n_launches * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / peaks["hbm_gbs"], "traffic": (traffic or {}).get("dram_bytes_per_launch"),
            "traffic_captured_at": (traffic or {}).get("commit"), "peak_source": peak_src,
            "ms_per_launch": ms / n_launches, "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launches_timed": n_launches}


KERNEL_NAMES = {  # layout -> (counts kernel, FAST kernel, EXACT kernel)
    "s16": ("counts_s16_kernel", "pheno_tc_kernel", "pheno_exact_s16d_kernel"),
    "group4": ("counts_g4_kernel", "pheno_lut_g4_kernel", "pheno_exact_g4_kernel"),
    "rows": ("counts_kernel", "pheno_lut_kernel", "pheno_exact_kernel"),
}


class Workload:
    """One BASELINE config resident on this rank's GPU: panel, vectors and the step function."""

    def __init__(self, args, name, rank, world, local, dev, stream):
        import torch
        import torch.distributed as dist
        from simu_b200 import capi
        self.torch, self.dist, self.capi = torch, dist, capi
        self.args, self.name, self.rank, self.world, self.dev, self.stream = args, name, rank, world, dev, stream
        w = self.w = WORKLOADS[name]
        n, m, mc, k = w["n"], w["m"], w["mc"], w["k"]
        self.ctx = ctx = capi.Context(n, device=local)
        ctx.set_stream(stream.cuda_stream)
        self.causal = causal = _causal_rows(w, rank)
        self.compact = args.layout in ("group4", "s16")
        self.layout_id = {"s16": capi.LAYOUT_SAMPLE16, "group4": capi.LAYOUT_GROUP4, "rows": capi.LAYOUT_ROWS}[args.layout]
        if self.compact:
            # what the drop-in CLI keeps in HBM: the causal rows only (the reference reads no others,
            # source.cc:852-863), compact and in causal order
            ctx.panel_alloc(mc, self.layout_id)
            if mc == m:
                ctx.panel_synth(PANEL_SEED, rank * m, m)
            else:
                ctx.panel_synth_rows(PANEL_SEED, causal.astype(np.int64) + rank * m)
            self.d_rows = None
        else:
            # the whole slice of the .bed resident, row-major; causal rows gathered through a row list
            ctx.panel_alloc(m)
            ctx.panel_synth(PANEL_SEED, rank * m, m)        # each GPU holds its own slice of the panel
            self.d_rows = torch.from_numpy(causal).to(dev) if mc != m else None
        self.rows_ptr = self.d_rows.data_ptr() if self.d_rows is not None else 0
        self.d_counts = torch.empty((mc, 4), dtype=torch.int32, device=dev)
        self.d_freq = torch.empty(mc, dtype=torch.float64, device=dev)
        ctx.find_freq_dev(self.rows_ptr, mc, self.d_counts.data_ptr(), self.d_freq.data_ptr())
        torch.cuda.synchronize()
        self.gcta = name == "config2"                       # --gcta-sigma: effects scaled by het^-1/2 between the passes
        r1, r2 = _raw_effects(w, rank)
        self.x1, self.x2 = _effects(w, self.d_freq.cpu().numpy(), rank) if self.gcta else (r1, r2)
        self.d_r1 = torch.from_numpy(r1).to(dev)
        self.d_r2 = torch.from_numpy(r2).to(dev) if k == 2 else None
        self.d_x1 = torch.from_numpy(self.x1).to(dev)
        self.d_x2 = torch.from_numpy(self.x2).to(dev) if k == 2 else None
        self.d_spow = torch.tensor([-1.0, -1.0], dtype=torch.float64, device=dev)
        self.d_bad = torch.zeros(1, dtype=torch.int64, device=dev)
        self.d_y = torch.zeros((k, n), dtype=torch.float64, device=dev)   # [trait][sample], contiguous for one all-reduce
        g = torch.Generator(device="cpu").manual_seed(5)
        self.d_noise = torch.randn((k, n), dtype=torch.float64, generator=g).to(dev)   # e: same draws on every rank
        self.d_misc = torch.zeros(8, dtype=torch.float64, device=dev)
        self.d_index = torch.empty(n, dtype=torch.int32, device=dev)
        self.d_is_case = torch.empty(n, dtype=torch.uint8, device=dev)
        self.max_ncon = n - int(np.floor(np.float32(0.1) * np.float32(n)))        # --k 0.1 (source.cc:1038-1039)
        self.mode = capi.MODE_EXACT if args.mode == "exact" else capi.MODE_FAST
        # the one exchange step: all-reduce over NVLink peer memory (simu_b200_peer_*), NCCL only if peer access is unavailable
        self.allreduce = "nccl"
        if world > 1 and args.allreduce in ("auto", "peer"):
            try:
                handles = [None] * world
                dist.all_gather_object(handles, ctx.peer_init(rank, world, k * n))
                ctx.peer_connect(handles)
                ok = torch.ones(1, device=dev)
            except capi.SimuB200Error as e:
                print(f"rank {rank}: peer all-reduce unavailable ({e}); using NCCL", file=sys.stderr)
                ok = torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            self.allreduce = "peer" if ok.item() > 0 else "nccl"
        # the scaling runs inside the step when the device-resident one-call form applies
        self.one_call = self.gcta and self.compact and self.mode == capi.MODE_FAST

    def step(self):
        ctx, w, k, n, mc = self.ctx, self.w, self.w["k"], self.w["n"], self.w["mc"]
        d_y = self.d_y
        if self.one_call:      # find_freq -> effect scaling (--gcta-sigma) -> find_pheno, no host round trip
            ctx.find_freq_pheno_dev(mc, self.d_r1.data_ptr(), self.d_r2.data_ptr() if k == 2 else 0, 0, 0, self.d_spow.data_ptr(), 1,
                                    self.d_counts.data_ptr(), self.d_freq.data_ptr(), self.d_x1.data_ptr(),
                                    self.d_x2.data_ptr() if k == 2 else 0, d_y[0].data_ptr(),
                                    d_y[1].data_ptr() if k == 2 else 0, self.d_bad.data_ptr())
        else:
            ctx.find_freq_dev(self.rows_ptr, mc, self.d_counts.data_ptr(), self.d_freq.data_ptr())
            ctx.find_pheno_dev(self.rows_ptr, mc, self.d_freq.data_ptr(), self.d_x1.data_ptr(),
                               self.d_x2.data_ptr() if k == 2 else 0, d_y[0].data_ptr(), d_y[1].data_ptr() if k == 2 else 0, self.mode)
        if self.world > 1:                                # the path's one exchange step
            if self.allreduce == "peer":
                ctx.peer_allreduce_dev(d_y.data_ptr(), k * n)
            else:
                self.dist.all_reduce(d_y)
        for t in range(k):
            ctx.apply_heritability_dev(HSQ[t], d_y[t].data_ptr(), n, self.d_noise[t].data_ptr(), self.d_misc[4 + 2 * t].data_ptr())
        if w["cc"]:      # default --ncas/--ncon: both pools fully labelled -> the order statistic decides
            if self.args.cc_sort:
                ctx.liability_order_dev(d_y[0].data_ptr(), n, self.d_index.data_ptr())
            else:
                ctx.liability_split_dev(d_y[0].data_ptr(), n, self.max_ncon, self.d_is_case.data_ptr())

    def measure(self, steps, warmup, sampler, prof_steps=20):
        """-> dict: value, ms_per_step, kernel_ms, rooflines, clocks, launches.  The timed region replays ONE
        CUDA graph of the step per iteration when the step could be captured (the launch-bound
        sparse configs are 10 launches of microseconds each); per-kernel times for the roofline come from a
        second, eager region with the library's event pairs around every launch."""
        torch, dist, ctx, w = self.torch, self.dist, self.ctx, self.w
        n, mc, k = w["n"], w["mc"], w["k"]
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        if steps <= 0:      # as many steps as fill ~60 ms, so that the 5 ms clock sampler sees the timed region
            t0 = time.perf_counter()
            self.step()
            torch.cuda.synchronize()
            steps = int(min(5000, max(10, 0.06 / max(time.perf_counter() - t0, 1e-6))))
            if self.world > 1:
                t = torch.tensor([steps], dtype=torch.int64, device=self.dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                steps = int(t.item())
        graph = None
        # (EXACT mode is timed eagerly: its SAMPLE16 kernel takes the effects as kernel parameters fetched from the device at call
        # time, which a stream capture does not allow -- under capture the library falls back to its slower table kernel)
        # N > 1: the exchange over NVLink peer memory is the library's own kernel with its epoch in device memory, so the whole
        # step -- exchange included -- is one graph on every rank (the NCCL fallback is launched eagerly)
        if (self.world == 1 or self.allreduce == "peer") and self.args.graph != "off" and self.mode == self.capi.MODE_FAST:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self.step()
                g.replay()
                torch.cuda.synchronize()
                graph = g
            except Exception as e:
                print(f"{self.name}: CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
                torch.cuda.synchronize()
                ctx.set_stream(self.stream.cuda_stream)
                self.step()
                torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record(self.stream)
        if graph is not None:
            for _ in range(steps):
                graph.replay()
        else:
            for _ in range(steps):
                self.step()
        e1.record(self.stream)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall1 = time.time()
        ms_total = e0.elapsed_time(e1)
        # kernels per step: counted over one eager step (a graph replay launches the same kernels)
        l0 = ctx.launches
        ctx.profile_enable(True)
        for _ in range(prof_steps):
            self.step()
        torch.cuda.synchronize()
        launches_per_step = (ctx.launches - l0) // prof_steps
        prof = {name: ctx.profile_read(i) for name, i in (("counts", 0), ("pheno_exact", 1), ("pheno_fast", 2),
                                                         ("pheno_fast_build", 3), ("pheno_fast_reduce", 4), ("cc_order", 6))}
        ctx.profile_enable(False)
        if self.world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        ms_step = ms_total / steps
        _, bytes_counts, bytes_pheno = alg_bytes(w)
        peaks, peak_src = _peaks()
        names = KERNEL_NAMES[self.args.layout]
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        groups = (mc + 15) // 16
        if self.args.layout == "s16" and groups < 8 * 2 * sms * 8 and groups * ((ctx.group_stride + 8191) // 8192) >= 2 * sms * 8:
            names = ("counts_s16_tma_kernel",) + names[1:]      # the library's own choice by shape (counts_s16.cu)
        fast = self.mode == self.capi.MODE_FAST
        miss_state, miss_entries = ctx.missing_index() if self.args.layout == "s16" else (0, 0)
        main_ms, main_n = prof["pheno_fast" if fast else "pheno_exact"]
        cnt_ms, cnt_n = prof["counts"]
        out = {
            "value": n * mc * k * self.world / (ms_step * 1e-3), "ms_per_step": ms_step, "steps": steps,
            "cuda_graph": graph is not None, "gpu_launches_per_step": int(launches_per_step),
            "bed_gbs": 2 * mc * ((n + 3) // 4) * self.world / (ms_step * 1e-3) / 1e9,
            "roofline": _roofline(names[1] if fast else names[2], bytes_pheno, main_ms, main_n, peaks, peak_src,
                                  _traffic(self.name, names[1] if fast else names[2], self.args.layout)),
            "roofline_counts": _roofline(names[0], bytes_counts, cnt_ms, cnt_n, peaks, peak_src,
                                         _traffic(self.name, names[0], self.args.layout)),
            "kernel_ms": {kname: (v[0] / v[1] if v[1] else None) for kname, v in prof.items()},
            "clocks": sampler.summary(t_wall0, t_wall1) if sampler else None,
            "checksum_device": float(self.d_y.abs().sum().item()),
            # find_pheno FAST on SAMPLE16 panels: the missing genotypes as a sparse list built once per panel (at its second
            # use) with the correction pass beside the main kernel, or as a dense plane of the MMA (include/simu_b200.h)
            "missing_genotypes": {1: f"sparse index, {miss_entries} entries", 2: "dense MMA plane", 0: "dense MMA plane (index not built)"}[miss_state],
        }
        if not fast and out["roofline"]:      # the ordered FP64 kernel is not an HBM kernel: say so next to the HBM fraction
            out["roofline"]["note"] = ("EXACT mode is bound by issue slots, the FP64 pipe and the shared-memory gather together "
                                       "(DESIGN.md 3 b1), not by HBM; the HBM fraction is reported for comparison with FAST only")
        return out

    def close(self):
        self.ctx.close()


def _traffic(workload, kernel, layout):
    """DRAM bytes per launch from a committed `ncu --set full` capture (profiles/ncu_traffic.json; each entry
    names the commit it was captured at), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t[f"{workload}:{layout}"][kernel]
    except Exception:
        return None


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    from simu_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; simu_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # bind this rank to the CPUs closest to its GPU BEFORE any pinned host memory is allocated: the staging
    # buffers of the e2e leg are then first touched on the GPU's NUMA node (8 ranks sharing one node's memory
    # controllers halve the per-GPU host->device rate; tools/h2d_probe.py measures both placements)
    affinity = "unchanged"
    if not args.no_affinity:
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
            affinity = f"NVML ideal CPUs of GPU {phys} ({len(os.sched_getaffinity(0))} of {os.cpu_count()})"
        except Exception as e:      # noqa: BLE001
            affinity = f"unchanged ({type(e).__name__})"
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local) if rank == 0 else None          # started before any panel is generated
    stream = torch.cuda.Stream(device=dev)           # one stream for torch ops, NCCL hand-off and our kernels
    torch.cuda.set_stream(stream)
    n, m, mc, k = w["n"], w["m"], w["mc"], w["k"]
    B, bytes_counts, bytes_pheno = alg_bytes(w)

    wl = Workload(args, args.workload, rank, world, local, dev, stream)
    ctx, mode, causal = wl.ctx, wl.mode, wl.causal
    res = wl.measure(args.steps, args.warmup, sampler)
    ms_step, value = res["ms_per_step"], res["value"]
    updates_step = n * mc * k * world
    x1, x2 = wl.x1, wl.x2
    # the e2e leg must see the effects the timed step used (scaled on the device inside the step for --gcta-sigma)
    if wl.one_call:
        x1 = wl.d_x1.cpu().numpy()
        x2 = wl.d_x2.cpu().numpy() if k == 2 else None
    y_check = res["checksum_device"]

    if args.no_e2e:
        if rank == 0:
            emit({"profiling_run": True, "ms_per_step": ms_step, "value": value, "kernel_ms": res["kernel_ms"],
                  "cuda_graph": res["cuda_graph"]})
        if sampler:
            sampler.stop()
        wl.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------ e2e through the host-buffer C ABI
    e2e_steps = max(2, min(args.steps, 5))
    host_rows = torch.empty((mc, B), dtype=torch.uint8, pin_memory=True)     # the causal rows as packed .bed rows
    if wl.compact:
        ctx.panel_get_rows_ptr(0, mc, host_rows.data_ptr(), B)
    else:
        panel_view = torch.as_tensor(_DevMem(ctx.panel_ptr, (m, ctx.row_stride)), device=dev)
        chunk = max(1, (1 << 30) // ctx.row_stride)
        for lo in range(0, mc, chunk):
            hi = min(mc, lo + chunk)
            idx = torch.from_numpy(causal[lo:hi].astype(np.int64)).to(dev)
            host_rows[lo:hi].copy_(panel_view.index_select(0, idx)[:, :B])
        torch.cuda.synchronize()
        del panel_view
    noise_h = wl.d_noise.cpu().numpy()
    wl.close()                                       # the e2e context stages its own panel
    ctx2 = capi.Context(n, device=local)
    ctx2.panel_alloc(mc, wl.layout_id if wl.compact else capi.LAYOUT_ROWS)
    e2e_y = None

    def e2e_step():
        nonlocal e2e_y
        ctx2.panel_put_rows_ptr(0, mc, host_rows.data_ptr(), B)            # pinned host -> HBM, side stream
        counts, freq = ctx2.find_freq(mc=mc)                                 # host outputs
        y1, y2 = ctx2.find_pheno(freq, x1, x2, mode=mode)                   # host x, freq in; host y out
        if world > 1:                                                        # partial sums -> one all-reduce
            t = torch.from_numpy(np.stack([y1, y2]) if k == 2 else y1[None]).to(dev)
            dist.all_reduce(t)
            t = t.cpu().numpy()
            y1, y2 = t[0], (t[1] if k == 2 else None)
        y1, _, _, _ = ctx2.apply_heritability(HSQ[0], y1, noise_h[0], None)
        if k == 2:
            y2, _, _, _ = ctx2.apply_heritability(HSQ[1], y2, noise_h[1], None)
        e2e_y = (y1, y2)
        return counts

    e2e_step()
    ctx2.sync()
    t0 = time.perf_counter()                       # staging alone (reported, not part of the timed e2e steps)
    ctx2.panel_put_rows_ptr(0, mc, host_rows.data_ptr(), B)
    ctx2.sync()
    stage_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx2.sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = mc * B + mc * 8 * (1 + k) + 2 * n * 8 * k
    d2h = mc * 24 + 2 * n * 8 * k + 16 * k
    e2e = {"value": updates_step / (e2e_ms * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "staging_ms": stage_ms, "staging_gbs": mc * B / stage_ms / 1e6,
           "host_cpu_affinity": affinity,
           "api": "simu_b200_panel_put_rows + find_freq + find_pheno + apply_heritability (host buffers)"}
    # the device-resident and host-buffer paths must agree (same inputs, same kernels)
    e2e_check = float(np.abs(e2e_y[0]).sum() + (np.abs(e2e_y[1]).sum() if k == 2 else 0.0))
    parity_ok = abs(e2e_check - y_check) <= 1e-9 * abs(y_check)
    ctx2.close()

    # ------------------------------------------------ CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        probe = 64
        rows_np = host_rows.numpy()
        t_probe, _ = cpu_reference_time(rows_np[:probe], n, k)
        s = int(max(probe, min(mc, 15.0 / max(t_probe / probe, 1e-7))))
        dt, kind = cpu_reference_time(rows_np[:s], n, k)
        cpu = {"value": n * s * k / dt, "unit": "updates/s", "cores": 1, "kind": kind,
               "sample": f"first {s} of {mc} causal SNP rows x {n} samples, one pass ({dt:.1f} s)",
               "host_cores": os.cpu_count()}
    del host_rows

    # ------------------------------------------------ the other BASELINE configs, same arm, fewer steps
    also = {}
    if not args.no_also:
        for name in ("config4", "config5", "config3", "config1"):
            if name == args.workload:
                continue
            torch.cuda.empty_cache()
            o = Workload(args, name, rank, world, local, dev, stream)
            r = o.measure(args.also_steps, 3, sampler, prof_steps=5)
            ow = WORKLOADS[name]
            also[name] = {"desc": ow["desc"], "n_samples": ow["n"], "n_causal_per_gpu": ow["mc"], "traits": ow["k"],
                          "parallelism": f"snp-shard x{world}, 1 all-reduce ({o.allreduce})" if world > 1 else "single GPU",
                          **{kk: r[kk] for kk in ("value", "ms_per_step", "steps", "cuda_graph", "gpu_launches_per_step", "bed_gbs",
                                                  "roofline", "roofline_counts", "kernel_ms", "clocks", "missing_genotypes")}}
            o.close()

    if rank == 0:
        compact = wl.compact
        panel_desc = {"s16": "causal rows only, compact, SAMPLE16 (16 rows sample-major, dosage-coded)",
                      "group4": "causal rows only, compact, GROUP4 (4 rows sample-interleaved)",
                      "rows": "whole .bed slice row-major + causal row list"}[args.layout]
        line = {
            "metric": "genotype_effect_updates_per_s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("s32-tensor-core/f64-recombine" if args.layout == "s16" else "f32-table/f64-accumulate") if mode == capi.MODE_FAST else "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "desc": w["desc"], "n_samples": n, "panel_rows_per_gpu": m,
                       "hbm_panel_rows_per_gpu": mc if compact else m,
                       "n_causal_per_gpu": mc, "traits": k, "mode": args.mode, "hbm_panel": panel_desc,
                       "step": "find_freq -> effect scaling on the device -> find_pheno -> apply_heritability per trait" if wl.one_call
                               else "find_freq -> find_pheno -> apply_heritability per trait" + (" -> liability split" if w["cc"] else ""),
                       "cuda_graph": res["cuda_graph"], "missing_genotypes": res["missing_genotypes"],
                       "resident_panel": "everything that depends on the panel alone is staging, not step: the SAMPLE16 transposition and, from "
                                         "the panel's second use, the index of its missing genotypes; allele counts are recomputed every step; "
                                         "the e2e leg re-stages the panel every step and therefore runs find_pheno with the dense missing-genotype plane",
                       "l2": f"inputs larger than L2 ({mc * B / 1e6:.0f} MB of causal rows per pass, 126 MB L2)",
                       "parallelism": f"snp-shard x{world}, 1 all-reduce ({wl.allreduce})" if world > 1 else "single GPU"},
            "bed_gbs": res["bed_gbs"],
            "roofline": res["roofline"], "roofline_counts": res["roofline_counts"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(res["gpu_launches_per_step"] * args.steps), "clocks": res["clocks"],
            "kernel_ms": res["kernel_ms"],
            "checksum": {"device": y_check, "e2e": e2e_check}, "parity_check": bool(parity_ok),
            "also": also,
        }
        emit(line)
    if sampler:
        sampler.stop()
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit(f"bench.py: device-resident and host-buffer results differ: {y_check!r} vs {e2e_check!r}")


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: route everything any library writes to fd 1
    (NCCL's version banner, torchrun notices) to stderr and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--layout", default="s16", choices=["s16", "group4", "rows"],
                    help="HBM panel: compact causal rows in SAMPLE16 layout (what the CLI stages; tensor-core FAST kernel), "
                         "in GROUP4 layout (lookup-table FAST kernel), or the whole .bed slice row-major with a causal row list")
    ap.add_argument("--graph", default="auto", choices=["auto", "off"],
                    help="single GPU: replay one CUDA graph of the step per timed iteration (auto) or launch eagerly (off)")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the process to the CPUs NVML names as closest to its GPU")
    ap.add_argument("--no-also", action="store_true", help="skip the `also` block (the other BASELINE configs)")
    ap.add_argument("--also-steps", type=int, default=0, help="steps per `also` workload (0: as many as fill ~60 ms)")
    ap.add_argument("--cc-sort", action="store_true",
                    help="--cc workloads: full liability argsort (needed when --ncas/--ncon select a subset) instead of the split")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: the partial-phenotype exchange as a one-shot all-reduce over NVLink peer memory, or NCCL")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
