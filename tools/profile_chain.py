#!/usr/bin/env python
"""Times scan_moments, reduce_finalize and the chain of a 200-scan stream separately (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_slam_b200 as g
import bench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
prec = sys.argv[2] if len(sys.argv) > 2 else "fp64"
scans, poses = bench.make_stream(S)
pl = g.LidarEvidencePipeline(n_bins=48, tau=0.1, precision=prec, n_points_cap=8192, lidar_origin=scans[0].lidar_origin)
res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans]
xis = np.stack([s.xi_body for s in scans]); pd = torch.from_numpy(poses).cuda()
jobs = pl.jobs_to_device([pl.make_job(res[k], xis[k]) for k in range(S)])
ms = g.create_empty_map_stats(48); iw = pl.create_lidar_iw_state(1e-3)
ctas = pl.plan_ctas(S, 7600); b = pl._buffers(S, ctas)
print("ctas/job", ctas)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("moments  ms", t(lambda: pl.launch_moments(jobs, None, S, ctas, res[0].layout, b["partials"])))
print("finalize ms", t(lambda: pl.launch_finalize(b["partials"], ctas, pl.rec, pl.rec * ctas, jobs, None, S, True, b["reduced"], b["stats"])))
for name, kw in (("chain iw+sigma", dict(iw_state=iw, use_sigma_meas=True)), ("chain iw", dict(iw_state=iw, use_sigma_meas=False)), ("chain plain", dict())):
    ms_ = t(lambda: pl.launch_chain(b["stats"], S, 1, ms, b["map_derived"], pd, pd, b["evidence"], **kw))
    print(f"{name:16s} ms {ms_:.3f}  us/scan {1e3 * ms_ / S:.1f}")
ms_ = t(lambda: pl.launch_chain(b["stats"], S, 1, ms, b["map_derived"], pd, None, b["evidence"]))
print(f"chain no-map-upd ms {ms_:.3f}  us/scan {1e3 * ms_ / S:.1f}")
ms_ = t(lambda: pl.run_stream(res, xis, pd, pd, ms, iw_state=iw, use_sigma_meas=True, jobs_dev=jobs, ctas=ctas))
print(f"whole stream pass ms {ms_:.3f}")
