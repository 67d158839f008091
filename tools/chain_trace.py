"""clock64 trace of the single-scan chain kernel (stage 2 of the host call): needs the trace build,
    make -C fl-slam_b200/csrc EXTRA=-DGCL_CHAIN_TRACE   (touch gcl_evidence.cu first),
which adds gcl_debug_chain_trace; prints the phase marks in microseconds from kernel start (profiles/r1_chain_kernel_single_scan_trace.md)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fl_slam_b200 as g
import bench
from fl_slam_b200 import _abi
S = 16
scans, poses = bench.make_stream(S)
pl = g.LidarEvidencePipeline(n_bins=48, tau=0.1, precision="fp64", n_points_cap=8192, lidar_origin=scans[0].lidar_origin)
B = 48
m = {k: np.zeros(v) for k, v in (("S_dir", (B, 3)), ("S_dir_scatter", (B, 3, 3)), ("N_dir", (B,)), ("N_pos", (B,)), ("sum_p", (B, 3)), ("sum_ppT", (B, 3, 3)))}
iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
names = {0: "start", 1: "staged in", 2: "r1 pass1", 3: "r1 sums", 4: "r1 svd+rot", 5: "r1 pair bar", 6: "r1 pre bar", 7: "r1 trans solve", 8: "r1 resid", 9: "r1 finish",
         10: "pre sigma", 11: "pre mder", 12: "pre sigma_p", 14: "r2 scatter", 15: "iw start", 16: "iw done", 17: "synced", 18: "map updated", 19: "flag"}
acc = np.zeros(32)
for k in range(S):
    pl.process_scan_host(scans[k].points, scans[k].timestamps, scans[k].weights, scans[k].scan_start_time, scans[k].scan_end_time,
                         scans[k].xi_body[None], poses[k][None], m, pose_update=poses[k], iw_state_host=iw_h, use_sigma_meas=True)
    torch.cuda.synchronize()
    out = (C.c_longlong * 32)()
    _abi.lib.gcl_debug_chain_trace(out)
    t = np.array(out[:], dtype=np.float64)
    if k >= 4:
        acc += (t - t[0])
        print("globaltimer ns", t[31] - t[30], "cycles", t[19] - t[0], "=> MHz", (t[19] - t[0]) / max(t[31] - t[30], 1) * 1e3)
acc /= (S - 4)
for i in sorted(names, key=lambda i: acc[i]):
    print(f"{names[i]:16s} {acc[i] / 1965.0:7.2f} us")
