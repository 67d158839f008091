#!/bin/bash
# One gpurun call for the end-of-milestone record: the whole GPU test suite, both bench arms, compute-sanitizer over the reduced subset
# (memcheck / racecheck / synccheck; tensor-core and CUDA-core fp32c kernels), then the ncu captures (tools/capture_profiles.sh).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest_exit=$?"; tail -3 gpurun_out/r2_gputests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref_exit=$?"
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench_exit=$?"
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_subset.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1)"
done
GCL_K1_TC=0 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_subset.py > gpurun_out/r2_sanitizer_racecheck_cudacore.log 2>&1
echo "racecheck(cuda-core) exit=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/r2_sanitizer_racecheck_cudacore.log | tail -1)"
bash tools/capture_profiles.sh
