#!/bin/bash
# One gpurun call for the end-of-milestone record: the whole GPU test suite, both bench arms, then the ncu captures
# (tools/capture_profiles.sh).  compute-sanitizer is closed on this pool (profiles/r2_sanitizer_unavailable.md): `tools/sanitize_subset.py` is the
# program to run under it elsewhere.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest_exit=$?"; tail -3 gpurun_out/r2_gputests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref_exit=$?"
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench_exit=$?"
bash tools/capture_profiles.sh
