#!/usr/bin/env python
"""Per-scan host time of the message -> evidence path (gcl_scan_submit_msg_host + evidence begin / end), lookahead 2.
    python tools/profile_e2e_msg.py [fp64|fp32c] [imu|xi] [n_rep]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_slam_b200 as g
from fl_slam_b200 import synth
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
mode = sys.argv[2] if len(sys.argv) > 2 else "imu"
n_rep = int(sys.argv[3]) if len(sys.argv) > 3 else 4
S, B = 48, 48
scans, poses = bench.make_stream(S)
pl = g.LidarEvidencePipeline(n_bins=B, tau=0.1, precision=prec, n_points_cap=8192, lidar_origin=scans[0].lidar_origin)
R_bl, t_bl = synth.rotation_from_rotvec(np.array([0.02, -0.01, 0.3])), scans[0].lidar_origin.copy()
msgs = []
for k, s in enumerate(scans):
    m = synth.make_pointcloud2(s, R_bl, t_bl, time_field="time_f32", time_base=s.scan_start_time - 0.1 * k - 5.0)
    buf = torch.from_numpy(m["data"]).pin_memory()
    imu = synth.make_imu_buffer(5.0 + 0.1 * k, 5.1 + 0.1 * k, 512, seed=k)
    imu.update(sigma=0.01, rotvec_start_WB=np.zeros((1, 3)), gyro_bias=np.zeros((1, 3)), accel_bias=np.zeros((1, 3)), gravity_W=np.array([0.0, 0.0, -9.81]))
    msgs.append((buf.numpy(), m, imu, buf))
xis = np.stack([s.xi_body for s in scans])


def run(n_rep):
    m = {k: np.zeros(v) for k, v in (("S_dir", (B, 3)), ("S_dir_scatter", (B, 3, 3)), ("N_dir", (B,)), ("N_pos", (B,)), ("sum_p", (B, 3)), ("sum_ppT", (B, 3, 3)))}
    iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
    acc = [0.0, 0.0, 0.0]
    def sub(k):
        kw = dict(imu=msgs[k][2]) if mode == "imu" else dict(xi_body=xis[k][None])
        pl.submit_scan_msg_host(k % 4, data=msgs[k][0], fields=msgs[k][1]["fields"], point_step=msgs[k][1]["point_step"], n_points=msgs[k][1]["n_points"],
                                header_stamp_sec=msgs[k][1]["header_stamp_sec"], R_base_lidar=R_bl, t_base_lidar=t_bl, **kw)
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
    torch.cuda.synchronize()
    tot = time.perf_counter() - t_all
    n = n_rep * S
    return 1e6 * tot / n, [1e6 * a / n for a in acc]


run(1)
tot, parts = run(n_rep)
print(f"{prec} {mode}: message path {tot:.1f} us/scan; host time in begin {parts[0]:.1f}, submit {parts[1]:.1f}, end {parts[2]:.1f}")
