#!/usr/bin/env python
"""Small driver for ncu: runs scan_moments (and the rest of the chain) a few times on one configuration.

    python tools/profile_k1.py --points 1048576 --bins 256 --precision fp32c --iters 6 [--chain]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1 << 20)
ap.add_argument("--cap", type=int, default=0)
ap.add_argument("--bins", type=int, default=256)
ap.add_argument("--precision", default="fp32c")
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--hyp", type=int, default=1)
ap.add_argument("--chain", action="store_true")
a = ap.parse_args()
cap = a.cap or a.points
sc = g.synth.make_scan(a.points, seed=1238, dense=a.points > 100000)
pl = g.LidarEvidencePipeline(n_bins=a.bins, tau=g.synth.DEFAULT_TAU, precision=a.precision, n_points_cap=cap, lidar_origin=sc.lidar_origin)
scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
ms = g.create_empty_map_stats(a.bins)
poses = torch.zeros(a.hyp, 6, dtype=torch.float64, device="cuda")
xis = np.repeat(sc.xi_body[None], a.hyp, axis=0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(a.iters):
    e0.record()
    if a.chain:
        pl.run_resident(scan, xis, poses, ms)
    else:
        _, n_sel = pl.selected_rows(a.points, cap)
        ctas = pl.plan_ctas(1, n_sel)
        pl.launch_moments(None, pl.make_job(scan, sc.xi_body), 1, ctas, scan.layout, pl._buffers(1, ctas)["partials"])
    e1.record()
    torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1) * 1e3:.1f} us")
