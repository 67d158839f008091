#!/bin/bash
# One-GPU record for profiles/: fp32c error table on the goldens, configs[0] report, configs[4] sweep, tensor-core report.
mkdir -p gpurun_out
timeout 600 python tools/fp32c_error_table.py > gpurun_out/r2_fp32c_error_table.md 2> gpurun_out/fp32c_err.err; echo "errtable_exit=$?"
timeout 300 python tools/cfg1_report.py > gpurun_out/r2_cfg1_report.txt 2>&1; echo "cfg1_exit=$?"
timeout 900 python tools/sweep_cfg5.py > gpurun_out/r2_cfg5_sweep_16Mx1024.md 2> gpurun_out/cfg5.err; echo "cfg5_exit=$?"
timeout 600 python tools/tc_report.py > gpurun_out/r2_tc_report.json 2> gpurun_out/tc_report.err; echo "tc_exit=$?"
tail -5 gpurun_out/r2_cfg1_report.txt; tail -12 gpurun_out/r2_cfg5_sweep_16Mx1024.md
