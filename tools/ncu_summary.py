#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a short markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xxx.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("inst_executed", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA-pipe instructions % of peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
    ("sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed", "tensor ops tf32 % of peak"),
    ("sm__ops_path_tensor_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed", "tensor ops bf16 % of peak"),
    ("smsp__sass_inst_executed_op_utcmma.sum", "UTCMMA instructions"),
    ("smsp__sass_inst_executed_op_tmem_ldt.sum", "LDTM instructions"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts of the tensor core (operand reads) % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts (all)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smem load bank conflicts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smem store bank conflicts"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local (spill) stores"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (MUFU/LDS)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio throttle"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none), read with `ncu -i ... --page raw --csv`\n")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}\n\n| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        print()


if __name__ == "__main__":
    main()
