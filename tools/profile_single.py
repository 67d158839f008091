#!/usr/bin/env python
"""One configs[1]-shaped scan, device resident, whole chain (scan_moments, reduce_finalize, chain) a few times: for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_slam_b200 as g
import bench
scans, poses = bench.make_stream(2)
pl = g.LidarEvidencePipeline(n_bins=48, tau=0.1, precision="fp64", n_points_cap=8192, lidar_origin=scans[0].lidar_origin)
res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans]
ms = g.create_empty_map_stats(48); iw = pl.create_lidar_iw_state(1e-3)
pd = torch.from_numpy(poses).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(6):
    k = it % 2
    e0.record()
    pl.run_resident(res[k], scans[k].xi_body[None], pd[k:k + 1], ms, pose_update=pd[k:k + 1], iw_state=iw, use_sigma_meas=True)
    e1.record(); torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1) * 1e3:.1f} us")

# per-kernel times (CUDA events, 50 launches each, back to back)
sc = res[0]; _, n_sel = pl.selected_rows(sc.n_raw, 8192); ctas = pl.plan_ctas(1, n_sel); b = pl._buffers(1, ctas)
job = pl.make_job(sc, scans[0].xi_body)
def t(fn, n=50):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
print("ctas", ctas)
print(f"scan_moments    {t(lambda: pl.launch_moments(None, job, 1, ctas, sc.layout, b['partials'])):.1f} us")
print(f"reduce_finalize {t(lambda: pl.launch_finalize(b['partials'], ctas, pl.rec, pl.rec * ctas, None, job, 1, True, b['reduced'], b['stats'])):.1f} us")
print(f"chain           {t(lambda: pl.launch_chain(b['stats'], 1, 1, ms, b['map_derived'], pd[:1], pd[:1], b['evidence'], iw, True, 0.99)):.1f} us")
print(f"chain (no iw)   {t(lambda: pl.launch_chain(b['stats'], 1, 1, ms, b['map_derived'], pd[:1], pd[:1], b['evidence'], None, False, 0.99)):.1f} us")
print(f"chain (ro map)  {t(lambda: pl.launch_chain(b['stats'], 1, 1, ms, b['map_derived'], pd[:1], None, b['evidence'], None, False, 0.99)):.1f} us")
for c in (37, 74):
    bb = pl._buffers(1, c)
    print(f"ctas {c}: scan_moments {t(lambda: pl.launch_moments(None, job, 1, c, sc.layout, bb['partials'])):.1f} us, reduce_finalize {t(lambda: pl.launch_finalize(bb['partials'], c, pl.rec, pl.rec * c, None, job, 1, True, bb['reduced'], bb['stats'])):.1f} us")
