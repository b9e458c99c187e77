#!/usr/bin/env python
"""Condense an .ncu-rep (full set) into the handful of numbers the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/x/render_full.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_registers", "occ limit regs"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem"), ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  of which shared %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "  LDS wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "  STS wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "  shared bank conflicts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "  L1 hits"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local load sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "local store sectors"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma) active %"),
    ("sm__inst_executed_pipe_tc.sum", "tensor-core instructions"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "FMA-heavy pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_xu.sum", "XU instructions"),
    ("sm__inst_executed_pipe_lsu.sum", "LSU instructions"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio_throttle"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not_selected"),
    ("smsp__average_warp_latency_issue_stalled_tex_throttle.ratio", "stall tex_throttle"),
    ("smsp__average_warp_latency_issue_stalled_membar.ratio", "stall membar"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append(f"### {d.get('Kernel Name', '?')}  (id {d.get('ID', '?')})\n")
        out.append("| metric | value | unit |\n|---|---|---|\n")
        for k, label in KEYS:
            if k in d and d[k] != "":
                out.append(f"| {label} (`{k}`) | {d[k]} | {u[k]} |\n")
        out.append("\n")
    text = "".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(text)
    print(text)


if __name__ == "__main__":
    main()
