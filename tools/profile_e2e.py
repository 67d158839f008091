#!/usr/bin/env python
"""Per-scan latency of the reference-facing host call (pinned inputs), with and without the Python layer.
GCL_HOST_TIMING=1 adds the per-stage breakdown printed by the library (serialised stages)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_slam_b200 as g
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
S = 64
scans, poses = bench.make_stream(S)
pl = g.LidarEvidencePipeline(n_bins=48, tau=0.1, precision=prec, n_points_cap=8192, lidar_origin=scans[0].lidar_origin)
pinned = []
for s in scans:
    p, t_, w_ = (torch.from_numpy(a).pin_memory() for a in (s.points, s.timestamps, s.weights))
    pinned.append((p.numpy(), t_.numpy(), w_.numpy(), p, t_, w_))
xis = np.stack([s.xi_body for s in scans])
B = 48
def run(n_rep):
    m = {k: np.zeros(v) for k, v in (("S_dir", (B, 3)), ("S_dir_scatter", (B, 3, 3)), ("N_dir", (B,)), ("N_pos", (B,)), ("sum_p", (B, 3)), ("sum_ppT", (B, 3, 3)))}
    iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
    t0 = time.perf_counter()
    for _ in range(n_rep):
        for k in range(S):
            pl.process_scan_host(pinned[k][0], pinned[k][1], pinned[k][2], scans[k].scan_start_time, scans[k].scan_end_time,
                                 xis[k][None], poses[k][None], m, pose_update=poses[k], iw_state_host=iw_h, use_sigma_meas=True)
    return 1e6 * (time.perf_counter() - t0) / (n_rep * S)
run(1)
print(f"{prec}: process_scan_host {run(4):.1f} us/scan")

# two-stage loop with per-call host timing
def run2(n_rep):
    m = {k: np.zeros(v) for k, v in (("S_dir", (B, 3)), ("S_dir_scatter", (B, 3, 3)), ("N_dir", (B,)), ("N_pos", (B,)), ("sum_p", (B, 3)), ("sum_ppT", (B, 3, 3)))}
    iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
    acc = [0.0, 0.0, 0.0]
    sub = lambda k: pl.submit_scan_host(k % 4, pinned[k][0], pinned[k][1], pinned[k][2], scans[k].scan_start_time, scans[k].scan_end_time, xis[k][None])
    t_all = time.perf_counter()
    for _ in range(n_rep):
        sub(0); sub(1)
        for k in range(S):
            t0 = time.perf_counter()
            pl.evidence_begin_host(k % 4, poses[k][None], m, pose_update=poses[k], iw_state_host=iw_h, use_sigma_meas=True)
            t1 = time.perf_counter()
            if k + 2 < S:
                sub(k + 2)
            t2 = time.perf_counter()
            pl.evidence_end_host()
            t3 = time.perf_counter()
            acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2
    tot = time.perf_counter() - t_all
    n = n_rep * S
    return 1e6 * tot / n, [1e6 * a / n for a in acc]
run2(1)
tot, parts = run2(4)
print(f"{prec}: two-stage {tot:.1f} us/scan; host time in begin {parts[0]:.1f}, submit {parts[1]:.1f}, end {parts[2]:.1f}")
