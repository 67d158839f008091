#!/bin/bash
# A/B builds: tools/build_variant.sh NAME "EXTRA NVCC FLAGS" file.cu [file.cu ...]
# compiles the named sources of fl-slam_b200/csrc with the extra flags, links them with the standard objects of the others into
# fl-slam_b200/libgcl_b200_NAME.so; run a tool against it with GCL_LIB=fl-slam_b200/libgcl_b200_NAME.so (fl-slam_b200/_abi.py).
set -e
NAME=$1; EXTRA=$2; shift 2
cd "$(dirname "$0")/../fl-slam_b200/csrc"
make -j8 > /dev/null
mkdir -p variants
OBJS=""
for f in gcl_api gcl_scan_fused gcl_scan_tc gcl_scan_moments gcl_finalize gcl_exchange gcl_pipe gcl_evidence gcl_imu gcl_ingest gcl_ops; do
  if [[ " $* " == *" $f.cu "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $EXTRA -c $f.cu -o variants/${NAME}_$f.o 2> variants/${NAME}_$f.o.ptxas.log || { cat variants/${NAME}_$f.o.ptxas.log; exit 1; }
    OBJS="$OBJS variants/${NAME}_$f.o"
  else
    OBJS="$OBJS $f.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgcl_b200_$NAME.so $OBJS -lcudart
echo "built fl-slam_b200/libgcl_b200_$NAME.so"
