#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a short text file for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for vals in raw[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== kernel: {name}")
        for i, h in enumerate(hdr):
            if h in KEYS or "issue_stalled" in h and "per_issue_active" in h:
                if vals[i] not in ("", "0"):
                    print(f"  {h:88s} {vals[i]:>18s} {units[i]}")
    src = page(rep, "source")
    for i, r in enumerate(src[:6]):
        if "Address" in r:
            h, start = r, i + 1
            break
    else:
        return
    ci = {k: i for i, k in enumerate(h)}
    rows = []
    for r in src[start:]:
        try:
            rows.append((int(r[ci["# Samples"]]), r))
        except Exception:
            pass
    tot = sum(s for s, _ in rows) or 1
    print(f"== top {top} instructions by stall samples (total {tot})")
    stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    for s, r in sorted(rows, key=lambda x: -x[0])[:top]:
        reasons = sorted(((int(r[ci[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
        why = ", ".join(f"{k}={v}" for v, k in reasons if v)
        print(f"  {s:7d} {100.0 * s / tot:5.1f}%  {r[ci['Source']][:70]:70s} {why}")


if __name__ == "__main__":
    main()
