#!/usr/bin/env python
"""clock64 phase trace of one CTA of the tensor-core scan_moments kernel (build with `make EXTRA=-DGCL_K1_TRACE`)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402

names = ["prologue", "row wait + front", "stage wait", "drain", "phase A", "reduction", "scale", "proxy fence + arrive", "-", "-", "-", "tail"]
for n, bins, cap in ((1 << 20, 256, 1 << 20), (1 << 20, 48, 1 << 20)):
    sc = g.synth.make_scan(n, seed=1238, dense=True)
    pl = g.LidarEvidencePipeline(n_bins=bins, tau=g.synth.DEFAULT_TAU, precision="fp32c", n_points_cap=cap, lidar_origin=sc.lidar_origin)
    scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
    ctas = pl.plan_ctas(1, n)
    job = pl.make_job(scan, sc.xi_body)
    part = pl._buffers(1, ctas)["partials"]
    for _ in range(3):
        pl.launch_moments(None, job, 1, ctas, scan.layout, part)
    torch.cuda.synchronize()
    out = (C.c_longlong * 24)()
    g._abi.lib.gcl_debug_tc_trace.restype = C.c_int
    rc = g._abi.lib.gcl_debug_tc_trace(out)
    rows = n // ctas
    for who, off in (("thread 0 (team 0)", 0), ("thread 256 (team 1)", 12)):
        v = [out[off + i] for i in range(12)]
        tot = sum(v)
        print(f"{n} x {bins}, {who}: total {tot} cycles for ~{rows} rows = {tot / rows:.1f} per point, {tot / (rows / 32):.0f} per sub-tile")
        print("   " + " | ".join(f"{nm} {100.0 * x / tot:.1f}% ({x / (rows / 32):.0f})" for nm, x in zip(names, v)))
