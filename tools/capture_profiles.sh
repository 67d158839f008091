#!/bin/bash
# One gpurun call: plain runs first (must exit 0), then the ncu launch lists and one --set full capture of the dominant kernel per
# precision.  Summaries are made from gpurun_out/*.ncu-rep / *.csv with tools/ncu_summary.py and committed under profiles/.
mkdir -p gpurun_out
for prec in fp64 fp32c; do
  CMD="python bench.py --precision $prec --steps 2 --warmup 3 --no-extra --no-cpu --no-e2e --scans-per-step 4"
  $CMD > gpurun_out/plain_$prec.json 2> gpurun_out/plain_$prec.err || { echo "plain run failed ($prec)"; tail -5 gpurun_out/plain_$prec.err; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_launches_bench_$prec.csv $CMD > /dev/null 2> gpurun_out/ncu_list_$prec.err
  ncu --set full --clock-control none --import-source on -k regex:scan_moments -s 8 -c 1 -o gpurun_out/r2_prof_scan_moments_$prec $CMD > /dev/null 2> gpurun_out/ncu_full_$prec.err
  ls -la gpurun_out/r2_prof_scan_moments_$prec.ncu-rep
done
tail -3 gpurun_out/r2_launches_bench_fp64.csv
