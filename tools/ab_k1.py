#!/usr/bin/env python
"""Timing of scan_moments alone on the named workloads (best of 3 x 30 launches, CUDA events) -- the harness used for A/B runs of
kernel variants (build the variant, run this, compare; GCL_K1_G forces the bins-per-thread tile).

    python tools/ab_k1.py [stream] [1M] [single] [hyp] [16M]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402

tau = g.synth.DEFAULT_TAU
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) / n)
    return best * 1e3  # us


def stream(precision, n_scans=200, pts=30000, cap=8192, bins=48, n_hyp=1):
    scans = [g.synth.make_scan(pts + 37 * (k % 11), seed=2000 + k) for k in range(n_scans)]
    pl = g.LidarEvidencePipeline(n_bins=bins, tau=tau, precision=precision, n_points_cap=cap, lidar_origin=scans[0].lidar_origin)
    res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans]
    jobs = pl.jobs_to_device([pl.make_job(res[k], scans[k].xi_body) for k in range(n_scans) for _ in range(n_hyp)])
    max_sel = max(pl.selected_rows(s.points.shape[0], cap)[1] for s in scans)
    ctas = pl.plan_ctas(n_scans * n_hyp, max_sel)
    b = pl._buffers(n_scans * n_hyp, ctas)
    us = timeit(lambda: pl.launch_moments(jobs, None, n_scans * n_hyp, ctas, res[0].layout, b["partials"]), 30)
    pairs = sum(pl.selected_rows(s.points.shape[0], cap)[1] for s in scans) * bins * n_hyp
    return {"us": us, "ctas_per_job": ctas, "Gpairs_per_s": pairs / us * 1e-3}


def single(precision, n, bins, cap=None, iters=30):
    cap = cap or n
    sc = g.synth.make_scan(n, seed=1238, dense=n > 100000)
    pl = g.LidarEvidencePipeline(n_bins=bins, tau=tau, precision=precision, n_points_cap=cap, lidar_origin=sc.lidar_origin)
    copies = [pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time) for _ in range((4 if n > 500000 else 32) if n > 100000 else 1)]
    _, n_sel = pl.selected_rows(n, cap)
    ctas = int(os.environ.get("GCL_AB_CTAS", "0")) or pl.plan_ctas(1, n_sel)      # GCL_AB_CTAS: override the planned grid
    part = pl._buffers(1, ctas)["partials"]
    jobs = [pl.make_job(c, sc.xi_body) for c in copies]
    state = {"i": 0}

    def fn():
        state["i"] += 1
        pl.launch_moments(None, jobs[state["i"] % len(jobs)], 1, ctas, copies[0].layout, part)
    us = timeit(fn, iters)
    return {"us": us, "ctas": ctas, "Gpairs_per_s": n_sel * bins / us * 1e-3}


out = {"GCL_K1_G": os.environ.get("GCL_K1_G")}
only = sys.argv[1:] or ["stream", "1M", "single", "hyp", "16M"]
for prec in (os.environ.get("GCL_AB_PREC", "fp64,fp32c").split(",")):
    if "stream" in only:
        out[f"stream200x30k_48_{prec}"] = stream(prec)
    if "1M" in only:
        out[f"1Mx256_{prec}"] = single(prec, 1 << 20, 256)
    if "shard8" in only:      # the shard one rank of eight sees of the 1 M x 256 cloud
        out[f"131072x256_{prec}"] = single(prec, 1 << 17, 256, iters=60)
    if "single" in only:
        out[f"8192x48_{prec}"] = single(prec, 8192, 48)
        out[f"30k_cap8192x48_{prec}"] = single(prec, 30000, 48, cap=8192)
    if "hyp" in only:
        out[f"64hyp_30kx48_{prec}"] = stream(prec, n_scans=1, pts=30000, cap=32768, bins=48, n_hyp=64)
    if "16M" in only:
        out[f"4Mx1024_{prec}"] = single(prec, 1 << 22, 1024, iters=3)
print(json.dumps(out))
