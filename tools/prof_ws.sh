CMD="python tools/ab_k1.py 1M"
GCL_K1_WS=1 ncu --set full --clock-control none --import-source on -k regex:ws64 -s 4 -c 1 -o gpurun_out/r2_prof_ws64 $CMD > /dev/null 2> gpurun_out/ncu_ws.err
ls -la gpurun_out/r2_prof_ws64.ncu-rep
