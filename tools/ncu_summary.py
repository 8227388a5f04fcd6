#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): `python tools/ncu_summary.py gpurun_out/prof.ncu-rep [pattern ...]`.
Prints, per profiled launch, the metrics whose names contain any of the patterns (a default roofline set if none)."""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ",
           "issue_stalled", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
           "smsp__cycles_active.avg", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    body = rows[1:]
    units = None
    if body and not body[0][0].isdigit():
        units, body = body[0], body[1:]
    for r in body:
        d = dict(zip(hdr, r))
        print("== launch", d.get("ID"), d.get("Kernel Name", "")[:100], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for i, h in enumerate(hdr):
            if any(p.strip() in h for p in pats):
                v = r[i]
                if v in ("", "n/a"):
                    continue
                print(f"   {h.split('.', 2)[-1] if h.count('.') > 3 else h:90s} {v} {units[i] if units else ''}")


if __name__ == "__main__":
    main()
