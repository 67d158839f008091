for v in 0 1 2 3; do
  cp tools/microbench/libgcl_dbg$v.so fl-slam_b200/libgcl_b200.so
  echo "=== GCL_TC_DEBUG=$v"
  timeout 200 python tools/tc_trace.py 2>&1 | tail -8
done
