#!/usr/bin/env python
"""clock64 split per role of one CTA of scan_moments_ws64_kernel on the dense cloud (1 M points x 256 bins, fp64): trace build
(tools/build_variant.sh wstrace "-DGCL_K1_TRACE" gcl_scan_fused.cu gcl_api.cu; GCL_LIB=fl-slam_b200/libgcl_b200_wstrace.so)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402

n, bins = 1 << 20, 256
sc = g.synth.make_scan(n, seed=1238, dense=True)
pl = g.LidarEvidencePipeline(n_bins=bins, tau=g.synth.DEFAULT_TAU, precision="fp64", n_points_cap=n, lidar_origin=sc.lidar_origin)
scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
ctas = pl.plan_ctas(1, n)
job = pl.make_job(scan, sc.xi_body)
part = pl._buffers(1, ctas)["partials"]
for _ in range(3):
    pl.launch_moments(None, job, 1, ctas, scan.layout, part)
torch.cuda.synchronize()
out = (C.c_longlong * 16)()
lib = C.CDLL(os.environ["GCL_LIB"])
lib.gcl_debug_ws_trace(out)
names = {0: "consumer warp 0: phase B", 1: "consumer warp 0: wait FULL", 4: "phase-A warp 0: wait EMPTY", 5: "phase-A warp 0: phase A",
         6: "phase-A warp 0: wait PROD", 8: "point warp: front end", 9: "point warp: wait PROD", 10: "point warp: scale"}
for i, nm in names.items():
    print(f"{nm:32s} {out[i] / 1.965e3:8.1f} us")
