#!/bin/bash
# One gpurun call for the tensor-core kernel: forced-kernel parity tests, time/error report against the CUDA-core kernel, clock64 trace.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_forced.py -m gpu -x -q > gpurun_out/tc_tests.log 2>&1; echo "pytest_exit=$?"; tail -15 gpurun_out/tc_tests.log
timeout 600 python tools/tc_report.py > gpurun_out/tc_report.json 2> gpurun_out/tc_report.err; echo "report_exit=$?"
python - <<P
import json
d=json.load(open("gpurun_out/tc_report.json"))
for k,v in d.items():
    for c,r in v.items():
        print(k[:12], c, round(r["fp32c_us"],1), round(r["fp64_us"],1), {a:float("%.2g"%b) for a,b in r["max_rel_err_vs_fp64"].items()})
P
if [ -f tools/microbench/libgcl_trace.so ]; then
  cp fl-slam_b200/libgcl_b200.so /tmp/keep.so; cp tools/microbench/libgcl_trace.so fl-slam_b200/libgcl_b200.so
  GCL_K1_TC=2 timeout 200 python tools/tc_trace.py 2>&1 | tail -12
  cp /tmp/keep.so fl-slam_b200/libgcl_b200.so
fi
