"""clock64 phase split of one CTA of scan_moments on the bench workload (stream, 200 scans x 7500 rows x 48 bins): needs the trace
build,  touch fl-slam_b200/csrc/gcl_scan_fused.cu; make -C fl-slam_b200/csrc EXTRA=-DGCL_K1_TRACE  (adds gcl_debug_k1_trace)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_slam_b200 as g
import bench
from fl_slam_b200 import _abi
prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
scans, poses = bench.make_stream(200)
pl = g.LidarEvidencePipeline(n_bins=48, tau=0.1, precision=prec, n_points_cap=8192, lidar_origin=scans[0].lidar_origin)
res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans]
xis = np.stack([s.xi_body for s in scans])
jobs = pl.jobs_to_device([pl.make_job(res[k], xis[k]) for k in range(200)])
ctas = pl.plan_ctas(200, 7500)
b = pl._buffers(200, ctas)
for _ in range(5):
    pl.launch_moments(jobs, None, 200, ctas, res[0].layout, b["partials"])
torch.cuda.synchronize()
out = (C.c_longlong * 16)()
_abi.lib.gcl_debug_k1_trace(out)
names = ["prologue (job, atlas, table, raw-weight sum)", "row wait + front end", "phase A", "scale", "phase B", "tail barrier", "flush", "certificates"]
tot = sum(out[i] for i in range(8))
for i, n in enumerate(names):
    print(f"{n:46s} {out[i] / 1.965e3:7.2f} us  {100.0 * out[i] / tot:5.1f} %")
print("CTA total", tot / 1.965e3, "us;", ctas, "CTAs per scan")
