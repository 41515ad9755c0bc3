#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md argues from.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv` is parsed and one block per profiled
launch is printed.  gpurun_out/ is scratch; the text this prints is what gets committed.
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(out.stdout)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of {rep.split('/')[-1]} ({len(rows) - 2} profiled launches)")
    print("# per-launch values; ncu replays each kernel cold-cache, so compare shares, not absolutes")
    for r in rows[2:]:
        print()
        print("kernel:", r[col["Kernel Name"]])
        for key in KEEP + extra:
            if key in col and r[col[key]] != "":
                print(f"  {key:88s} {r[col[key]]} {units[col[key]]}")
        tens = [h for h in hdr if "tensor" in h and h.endswith("pct_of_peak_sustained_active") and h not in KEEP]
        for h in tens:
            if r[col[h]] not in ("", "0"):
                print(f"  {h:88s} {r[col[h]]} {units[col[h]]}")


if __name__ == "__main__":
    main()
