#!/usr/bin/env python
"""Does the upload of a 1 M-point scan (40 MB of page-locked float64 rows, copy engine) run slower while scan_moments occupies the SMs?
H2D rate alone, beside the fp64 kernel, beside the fp32c kernel; and the kernels' own times beside the upload.

    python tools/h2d_probe.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402

n, B = 1 << 20, 256
sc = g.synth.make_scan(n, seed=1238, dense=True)
pin = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (sc.points, sc.timestamps, sc.weights)]
dst = [torch.empty_like(t, device="cuda") for t in pin]
nbytes = sum(t.numel() * t.element_size() for t in pin)
cs, ks = torch.cuda.Stream(), torch.cuda.Stream()
out = {}


def h2d_rate(n_rep=24):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record()
        for _ in range(n_rep):
            for d, p in zip(dst, pin):
                d.copy_(p, non_blocking=True)
        e1.record()
    return e0, e1, n_rep


for prec in (None, "fp64", "fp32c"):
    if prec is None:
        for _ in range(2):
            e0, e1, r = h2d_rate()
            torch.cuda.synchronize()
        out["h2d_alone_gbs"] = nbytes * r / (e0.elapsed_time(e1) * 1e-3) * 1e-9
        continue
    pl = g.LidarEvidencePipeline(n_bins=B, tau=g.synth.DEFAULT_TAU, precision=prec, n_points_cap=n, lidar_origin=sc.lidar_origin)
    scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
    ctas = pl.plan_ctas(1, n) - 2
    job = pl.make_job(scan, sc.xi_body)
    part = pl._buffers(1, ctas)["partials"]
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_k = 40 if prec == "fp64" else 120

    def kernels():
        with torch.cuda.stream(ks):
            k0.record()
            for _ in range(n_k):
                pl.launch_moments(None, job, 1, ctas, scan.layout, part)
            k1.record()
    kernels(); torch.cuda.synchronize()
    kernels(); torch.cuda.synchronize()
    out[f"{prec}_kernel_alone_us"] = k0.elapsed_time(k1) / n_k * 1e3
    kernels()
    e0, e1, r = h2d_rate()
    torch.cuda.synchronize()
    out[f"h2d_beside_{prec}_gbs"] = nbytes * r / (e0.elapsed_time(e1) * 1e-3) * 1e-9
    out[f"{prec}_kernel_beside_h2d_us"] = k0.elapsed_time(k1) / n_k * 1e3
print(json.dumps(out, indent=1))
