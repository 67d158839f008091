#!/bin/bash
# One gpurun call: ncu --set full capture of the configs[1] stream launch of scan_moments (200 scans x 7500 rows x 48 bins, grid (5, 200)) out
# of the bench command itself.  The 48-bin fp64 kernel is also what the hypothesis batches launch, so the stream's launch is found by its
# grid in a launch list taken first.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --scans-per-step 2"
$CMD > gpurun_out/stream_plain.json 2> gpurun_out/stream_plain.err || { echo "plain run failed"; tail -5 gpurun_out/stream_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"scan_moments_fused_kernel<double, .int.3," --csv --log-file gpurun_out/r2_launches_48bins.csv $CMD > /dev/null 2> gpurun_out/ncu_list_stream.err
SKIP=$(python - <<P
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_48bins.csv")) if len(r) > 8 and r[0].isdigit()]
idx = [i for i, r in enumerate(rows) if r[8].replace(" ", "") == "(5,200,1)"]
print(idx[min(2, len(idx) - 1)] if idx else -1)
P
)
echo "stream launch index among the 48-bin fp64 launches: $SKIP"
[ "$SKIP" -ge 0 ] || { grep -o '([0-9, ]*)' gpurun_out/r2_launches_48bins.csv | sort | uniq -c | head; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"scan_moments_fused_kernel<double, .int.3," -s $SKIP -c 1 -o gpurun_out/r2_prof_scan_moments_stream $CMD > /dev/null 2> gpurun_out/ncu_full_stream.err
ls -la gpurun_out/r2_prof_scan_moments_stream.ncu-rep
