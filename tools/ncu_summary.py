#!/usr/bin/env python
"""Condense an ncu report (--set full) into the few numbers DESIGN.md / bench.py quote.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_op_read_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none summary of {rep}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n## kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k in KEYS:
            if k in d:
                print(f"{k:95s} {d[k]:>18s} {units[hdr.index(k)]}")
        rd, wr = float(d.get("dram__bytes_read.sum", 0) or 0), float(d.get("dram__bytes_write.sum", 0) or 0)
        u_rd, u_wr = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        print(f"{'traffic = dram read + write':95s} {rd:.3f} {u_rd} + {wr:.3f} {u_wr}")


if __name__ == "__main__":
    main()
