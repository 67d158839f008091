#!/usr/bin/env python
"""clock64 phase split of one CTA of scan_moments_fused_kernel on the dense cloud (1 M points x 256 bins, fp64 by default): trace build,
make -C fl-slam_b200/csrc EXTRA=-DGCL_K1_TRACE."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
n, bins = (int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20), 256
n_ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sc = g.synth.make_scan(n, seed=1238, dense=True)
pl = g.LidarEvidencePipeline(n_bins=bins, tau=g.synth.DEFAULT_TAU, precision=prec, n_points_cap=n, lidar_origin=sc.lidar_origin)
scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
ctas = n_ctas or pl.plan_ctas(1, n)
job = pl.make_job(scan, sc.xi_body)
part = pl._buffers(1, ctas)["partials"]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    e0.record()
    pl.launch_moments(None, job, 1, ctas, scan.layout, part)
    e1.record()
torch.cuda.synchronize()
print(f"launch (events, trace build): {e0.elapsed_time(e1) * 1e3:.1f} us for {n} rows on {ctas} CTAs")
out = (C.c_longlong * 16)()
g._abi.lib.gcl_debug_k1_trace(out)
names = ["prologue (job, atlas, table, raw-weight sum)", "row wait + front end", "phase A", "scale", "phase B", "tail barrier", "flush", "certificates"]
tot = sum(out[i] for i in range(8))
for i, nm in enumerate(names):
    print(f"{nm:46s} {out[i] / 1.965e3:8.2f} us  {100.0 * out[i] / tot:5.1f} %")
print("CTA total", tot / 1.965e3, "us;", ctas, "CTAs")
