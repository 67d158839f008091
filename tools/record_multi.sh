#!/bin/bash
# Multi-GPU record (run under `gpurun --gpus N`): the sharded-path parity tests, the bench line at N ranks, and (N = 8) configs[4]
# (16 M points x 1024 bins) sharded by points in both precisions.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_gpu_multi_tests_n$N.log 2>&1; echo "multi_tests_exit=$?"; tail -3 gpurun_out/r2_gpu_multi_tests_n$N.log
timeout 900 $TR bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench_exit=$?"
if [ "$N" = "8" ]; then
  for prec in fp64 fp32c; do
    timeout 900 $TR bench.py --gpus $N --precision $prec --head-points 16777216 --head-bins 1024 --scans-per-step 4 --steps 3 --no-extra --no-cpu --no-e2e \
      > gpurun_out/r2_bench_cfg5_n${N}_$prec.json 2> gpurun_out/r2_bench_cfg5_n${N}_$prec.err; echo "cfg5 $prec exit=$?"
  done
fi
python - <<P
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench*_n$N*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.3e"%d["value"], d["details"]["ms_per_scan"], d["details"]["scan_moments_ms"], d["details"]["rest_of_scan_ms"], d["parity"]["ranks_bit_identical"], max(d["parity"]["vs_unsharded_max_rel"].values()))
    for k,v in d.get("extra",{}).items():
        print("   ",k,{a:v[a] for a in v if a in ("ms_per_scan","scan_moments_ms","rest_of_scan_ms","ms_per_batch","points_per_sec","ms_per_step")})
P
