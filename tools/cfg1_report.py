#!/usr/bin/env python
"""BASELINE.json configs[0]: one synthetic scan, 8192 points, 48 bins, 1 hypothesis -- GPU latency (device-resident chain and the
host call), the CPU oracle on the same arrays, and the errors between them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_slam_b200 as g
from oracle import gc_oracle as O

n, cap, B, tau = 8192, 8192, 48, g.synth.DEFAULT_TAU
sc = g.synth.make_scan(n, seed=1234 + 1)
bins = O.create_fibonacci_atlas(B)
def oracle():
    pb = O.point_budget_resample(sc.points, sc.timestamps, sc.weights, sc.ring, sc.tag, cap)
    dk = O.deskew_constant_twist(pb["points"], pb["timestamps"], pb["weights"], sc.scan_start_time, sc.scan_end_time, sc.xi_body)
    raw = O.fused_scan_raw_sums(dk["points"], dk["weights"], bins, tau, sc.lidar_origin)
    fin = O.scan_bin_finalize(raw)
    return raw, fin
raw, fin = oracle()
Rs, ts = g.synth.known_map_transform(0)
ms = g.synth.map_stats_from_scan_sums(raw, Rs, ts, seed=0)
pose = np.zeros(6)
t0 = time.perf_counter(); reps = 10
for _ in range(reps):
    raw, fin = oracle()
    md = O.compute_map_derived_stats(ms["S_dir"], ms["N_dir"], ms["N_pos"], ms["sum_p"], ms["sum_ppT"])
    mf = O.matrix_fisher_rotation_evidence(pose, fin["s_dir"], fin["S_dir_scatter"], fin["N"], ms["S_dir"], ms["S_dir_scatter"], ms["N_dir"])
cpu_s = (time.perf_counter() - t0) / reps
print(f"CPU oracle: {1e3 * cpu_s:.2f} ms/scan = {n / cpu_s:.3e} points/s (NumPy, {os.cpu_count()} cores visible)")
for prec in ("fp64", "fp32c"):
    pl = g.LidarEvidencePipeline(n_bins=B, tau=tau, precision=prec, n_points_cap=cap, lidar_origin=sc.lidar_origin, bin_directions=bins)
    scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
    msd = g.MapBinStats.from_arrays(ms)
    pd = torch.zeros(1, 6, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5): ev = pl.run_resident(scan, sc.xi_body[None], pd, msd)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): ev = pl.run_resident(scan, sc.xi_body[None], pd, msd)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 200
    pin = [torch.from_numpy(a).pin_memory().numpy() for a in (sc.points, sc.timestamps, sc.weights)]
    for _ in range(5): evh = pl.process_scan_host(*pin, sc.scan_start_time, sc.scan_end_time, sc.xi_body[None], pose[None], ms)
    t0 = time.perf_counter()
    for _ in range(200): evh = pl.process_scan_host(*pin, sc.scan_start_time, sc.scan_end_time, sc.xi_body[None], pose[None], ms)
    ush = 1e6 * (time.perf_counter() - t0) / 200
    sb = ev.scan_bins(0)
    err = max(float(np.max(np.abs(getattr(sb, k).cpu().numpy() - ref)) / np.max(np.abs(ref)))
              for k, ref in (("N", fin["N"]), ("s_dir", fin["s_dir"]), ("S_dir_scatter", fin["S_dir_scatter"]), ("sum_p", raw["sum_p"]), ("sum_ppT", raw["sum_ppT"])))
    R = ev.evidence[0, :9].cpu().numpy().reshape(3, 3)
    ang = float(np.linalg.norm(R - mf["R_mf"]) / np.sqrt(2.0))   # = 2 sin(theta/2): the arccos form bottoms out at 1.5e-8
    print(f"{prec}: device-resident chain {us:.1f} us/scan = {n / us * 1e6:.3e} points/s, {n * B / us * 1e6:.3e} pairs/s; host call {ush:.1f} us/scan = "
          f"{n / ush * 1e6:.3e} points/s; max rel err raw sums {err:.2e}; R_mf geodesic {ang:.2e} rad")
