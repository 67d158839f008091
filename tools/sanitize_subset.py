#!/usr/bin/env python
"""A reduced pass over the kernels with non-trivial synchronisation, for compute-sanitizer (one tool per run):

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_subset.py

Covers: scan_moments fp64 (register-tile kernel, flush path, shared-memory aliasing of the flush staging on the e-tile), scan_moments
fp32c on the tensor cores (mbarriers, tcgen05, two stages, drain) and on the CUDA cores (GCL_K1_TC=0 in a second run), reduce_finalize,
the fused exchange kernel with one rank (ticket, flags), chain_kernel (named barriers, polled ticket of the host path), the staged
stream chain (stream_sigma_chain_kernel), PointCloud2 ingest (ticket), IMU preintegration.  Sizes are small: the sanitizer slows
kernels down by 10-100x."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402
from fl_slam_b200 import sharding as S  # noqa: E402

tau = g.synth.DEFAULT_TAU
sc = g.synth.make_scan(5000, seed=1)
for precision in ("fp64", "fp32c"):
    for B, cap in ((48, 2048), (256, 8192)):
        pl = g.LidarEvidencePipeline(n_bins=B, tau=tau, precision=precision, n_points_cap=cap, lidar_origin=sc.lidar_origin)
        scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
        ms = g.create_empty_map_stats(B)
        ev = pl.run_resident(scan, sc.xi_body[None], torch.zeros(1, 6, dtype=torch.float64, device="cuda"), ms)
        sh = S.PointShardedEvidence(pl, 0, 1, exchange="peer")
        ev2 = sh.run(scan, sc.points.shape[0], sc.xi_body[None], torch.zeros(1, 6, dtype=torch.float64, device="cuda"), ms)
        torch.cuda.synchronize()
        assert torch.isfinite(ev.stats).all() and torch.isfinite(ev2.stats).all()
        sh.px.close()
# stream chain (staged) + host path (self-staging chain kernel with the polled ticket) + message path (ingest ticket, IMU)
S_, B, cap = 6, 48, 2048
scans = [g.synth.make_scan(4000 + 31 * k, seed=10 + k, t0=g.synth.T0_EPOCH + 0.1 * k) for k in range(S_)]
poses = np.stack([np.array([0.05 * k, 0.01 * k, 0.0, 0.0, 0.0, 0.01 * k]) for k in range(S_)])
pl = g.LidarEvidencePipeline(n_bins=B, tau=tau, precision="fp64", n_points_cap=cap, lidar_origin=scans[0].lidar_origin)
res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans]
ms, iw = g.create_empty_map_stats(B), pl.create_lidar_iw_state(1e-3)
pd = torch.from_numpy(poses).cuda()
pl.run_stream(res, np.stack([s.xi_body for s in scans]), pd, pd, ms, iw_state=iw, use_sigma_meas=True)
torch.cuda.synchronize()
m = {k: np.zeros(v) for k, v in (("S_dir", (B, 3)), ("S_dir_scatter", (B, 3, 3)), ("N_dir", (B,)), ("N_pos", (B,)), ("sum_p", (B, 3)), ("sum_ppT", (B, 3, 3)))}
iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
for k in range(3):
    s = scans[k]
    pl.process_scan_host(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time, s.xi_body[None], poses[k][None], m,
                         pose_update=poses[k], iw_state_host=iw_h, use_sigma_meas=True)
R_bl, t_bl = g.synth.rotation_from_rotvec(np.array([0.02, -0.01, 0.3])), scans[0].lidar_origin.copy()
msg = g.synth.make_pointcloud2(scans[0], R_bl, t_bl, time_field="time_f32", time_base=scans[0].scan_start_time - 5.0)
imu = g.synth.make_imu_buffer(5.0, 5.1, 512, seed=0)
imu.update(sigma=0.01, rotvec_start_WB=np.zeros((1, 3)), gyro_bias=np.zeros((1, 3)), accel_bias=np.zeros((1, 3)), gravity_W=np.array([0.0, 0.0, -9.81]))
pl.submit_scan_msg_host(0, data=msg["data"], fields=msg["fields"], point_step=msg["point_step"], n_points=msg["n_points"],
                        header_stamp_sec=msg["header_stamp_sec"], R_base_lidar=R_bl, t_base_lidar=t_bl, imu=imu)
pl.evidence_host(0, poses[0][None], m)
torch.cuda.synchronize()
print("sanitize subset done")
