#!/usr/bin/env python
"""bench.py -- throughput of the LiDAR evidence hot path on B200 (contract in the task brief, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp64|fp32c] [--no-extra] [--no-cpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload = BASELINE.json configs[3], the configuration the north star's HBM / scaling target is quoted on: one dense
synthetic cloud of 1 048 576 points (n_points_cap = n => stride 1), 256 bins, 1 hypothesis, through resample -> deskew -> soft
assign -> moment match -> Matrix Fisher -> planar translation.  One "step" = 64 scans; the cloud is rotated through enough
copies that no scan re-reads an L2-resident input.  At N > 1 the POINTS are sharded over the ranks (strong scaling) with one
exchange of the per-bin record per scan through the fused NVLink kernel (csrc/gcl_exchange.cu); the sharded result is compared
with the unsharded one and bit for bit across ranks inside the run (`parity`).  `value` = points per second with the inputs
resident in HBM (fp64 arithmetic; the fp32c mode is in `extra`); `e2e` = the same scans with every scan's rows crossing PCIe from
page-locked host arrays and every scan's records read back; `roofline` is for the dominant kernel, scan_moments.
`extra` carries the same cloud in fp32c mode, the configs[1] scan stream (200 scans x ~30k points, map + IW state carried; at
N > 1 it runs as N independent replicas: scan k+1 needs scan k's map) with its per-scan host-call latencies, and
the configs[2] 64-hypothesis batch sharded by hypothesis.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SCANS, PTS_PER_SCAN, CAP, N_BINS = 200, 30000, 8192, 48          # configs[1] (side measurement)
HEAD_POINTS, HEAD_BINS, SCANS_PER_STEP = 1 << 20, 256, 64           # configs[3] (headline)
REF_SAMPLE_POINTS = 1 << 16                                         # points of the same cloud one reference-arm step processes
# dram__bytes_read.sum + dram__bytes_write.sum of ONE scan_moments launch of this workload, from the ncu --set full
# capture summarised in profiles/ (None until captured)
TRAFFIC_BYTES_PER_LAUNCH = {"fp64": 240.096e6 + 8.576e6, "fp32c": None}
# the same for one launch of the headline workload (1M x 256, one GPU); filled from profiles/r2_*.md once captured
HEAD_TRAFFIC_BYTES_PER_LAUNCH = {"fp64": 42032896, "fp32c": 21059584}   # dram__bytes_read.sum + dram__bytes_write.sum
HEAD_TRAFFIC_SOURCE = ("profiles/r2_scan_moments_fp64_1Mx256.md, r2_scan_moments_fp32c_1Mx256.md (ncu --set full capture of one scan_moments launch of this "
                       "workload: 1.002x / 1.004x the algorithmic bytes)")
TRAFFIC_SOURCE = ("profiles/r2_scan_moments_fp64_stream200x30k_48bins.md (dram__bytes_read.sum + dram__bytes_write.sum of one launch, captured out of "
                  "this command with tools/capture_stream_profile.sh; the scans are resident in the reference's layout -- all ~30 000 raw rows of 24 + 8 + 8 "
                  "bytes -- and PointBudgetResample keeps every 4th: DRAM serves 64-byte requests, so the stride-4 gather of 24-byte rows at a 96-byte "
                  "pitch and of 8-byte times at a 32-byte pitch fetches every raw byte, 30 000 x 40 B = 1.2 MB per scan against 0.54 MB algorithmic, "
                  "i.e. ~2.3x; the message path decodes straight into compact rows (GCL_JOB_COMPACT_ROWS) and does not pay it)")


# FMA-issue ceilings of scan_moments (rates measured with tools/microbench/ffma_rate.cu):
# FMA-pipe lane operations per (point, bin) pair = 19 accumulate + 3 similarity + softmax bookkeeping (+ the fp64 exp): fp64 phase A is 12
# FP64-pipe operations per pair since the 1/1024-octave exp table (round 2; it was 15 in round 1, i.e. 34 in all).  The default fp64 kernel
# at 129..256 bins issues part of them as DMMA (the dot as a K = 4 product with one padding lane, the 19 moments padded to 24 rows: 37 lane
# operations per pair on the pipe); the ceiling keeps the 31 USEFUL operations, so the padding counts against the kernel
FMA_LANES_PER_CLK_SM = {"fp32c": 128.0, "fp64": 64.0}
LANE_OPS_PER_PAIR = {"fp32c": 24.0, "fp64": 31.0}
N_SM = 148


def pair_ceiling(precision, sm_mhz):
    """pairs/s at which the FP32 (FP64) FMA pipe of all SMs would be saturated by scan_moments' own instruction mix."""
    return FMA_LANES_PER_CLK_SM[precision] / LANE_OPS_PER_PAIR[precision] * N_SM * sm_mhz * 1e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_stream(n_scans, seed0=1234 + 2):
    """configs[1]: n = 30000 +- 2 % per scan, 10 Hz sweeps, a slow known trajectory for the map update."""
    from fl_slam_b200 import synth
    rng = np.random.default_rng(seed0)
    scans, poses = [], []
    for k in range(n_scans):
        n = int(round(PTS_PER_SCAN * (1 + 0.02 * rng.uniform(-1, 1))))
        scans.append(synth.make_scan(n, seed=seed0 + 1 + k, t0=synth.T0_EPOCH + 0.1 * k))
        poses.append(np.array([0.05 * k, 0.01 * k, 0.0, 0.0, 0.0, 0.008 * k]))
    return scans, np.stack(poses)


def head_config(tau, precision="fp64"):
    """The `config` both arms print (the reference arm times a bounded sample of the same cloud; it says so in cpu_baseline)."""
    label = "configs[3]" if (HEAD_POINTS, HEAD_BINS) == (1 << 20, 256) else "configs[4]" if (HEAD_POINTS, HEAD_BINS) == (1 << 24, 1024) else "custom"
    return {"workload": f"{label}: dense synthetic cloud, {HEAD_POINTS} points/scan, {HEAD_BINS} bins, 1 hypothesis, n_points_cap = n "
                        "(stride 1); resample -> deskew -> soft assign -> moment match -> Matrix Fisher -> planar translation",
            "n_points": HEAD_POINTS, "n_bins": HEAD_BINS, "n_hyp": 1, "tau": tau, "precision": precision}


def _load_by_path(name, rel):
    """A module of the package loaded by file path: the reference arm must not import the package (whose __init__ loads the CUDA
    library)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _synth():
    return sys.modules.get("fl_slam_b200.synth") or sys.modules.get("_gcl_synth") or _load_by_path("_gcl_synth", "fl-slam_b200/synth.py")


def cpu_head_rate(n_sample, tau, min_seconds=0.0, return_result=False):
    """The NumPy restatement of the reference (JAX unavailable) on a bounded sample of the SAME cloud: the first n_sample points
    of the 1M-point cloud (a contiguous arc of the sweep) against all 256 bins -- resample, deskew, soft assign, moment match
    (chunked so that no (N, B) array exceeds 64 M elements), finalisation, Matrix Fisher, planar translation."""
    from oracle import gc_oracle as O
    synth = _synth()
    sc = synth.make_scan(HEAD_POINTS, seed=1234 + 4, dense=True)
    pts, ts, w = sc.points[:n_sample], sc.timestamps[:n_sample], sc.weights[:n_sample]
    bins = O.create_fibonacci_atlas(HEAD_BINS)
    dt, passes, out = 0.0, 0, None
    while passes == 0 or dt < min_seconds:
        t0 = time.perf_counter()
        pb = O.point_budget_resample(pts, ts, w, sc.ring[:n_sample], sc.tag[:n_sample], n_sample)
        dk = O.deskew_constant_twist(pb["points"], pb["timestamps"], pb["weights"], sc.scan_start_time, sc.scan_end_time, sc.xi_body)
        raw = O.fused_scan_raw_sums(dk["points"], dk["weights"], bins, tau, sc.lidar_origin)
        fin = O.scan_bin_finalize(raw)
        Rs, tstar = synth.known_map_transform(4)
        ms = synth.map_stats_from_scan_sums(raw, Rs, tstar, seed=4)
        md = O.compute_map_derived_stats(ms["S_dir"], ms["N_dir"], ms["N_pos"], ms["sum_p"], ms["sum_ppT"])
        mf = O.matrix_fisher_rotation_evidence(np.zeros(6), fin["s_dir"], fin["S_dir_scatter"], fin["N"], ms["S_dir"], ms["S_dir_scatter"], ms["N_dir"])
        pt = O.planar_translation_evidence(np.zeros(6), fin["p_bar"], fin["Sigma_p"], fin["N"], md["centroid"], md["Sigma_c"], ms["N_pos"],
                                           ms["S_dir_scatter"], ms["N_dir"], mf["R_mf"])
        dt += time.perf_counter() - t0
        passes += 1
        out = (sc, raw, fin, ms, mf, pt)
    rate = passes * n_sample / dt
    return (rate, dt, passes, out) if return_result else (rate, dt, passes)


def cpu_single_thread_rate(n_sample, tau):
    """The same sample with BLAS / OpenMP pools limited to one thread (SURVEY.md 8d asks for both figures)."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        return None
    with threadpool_limits(limits=1):
        return cpu_head_rate(n_sample, tau)[0]


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(i.get("num_threads", 1)) for i in threadpool_info()] + [1])
    except Exception:
        return 1


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path = the NumPy restatement (the JAX implementation cannot
    run here: no JAX, no network), all host threads NumPy/BLAS will use, one bounded sample of the headline cloud per step.  Imports
    nothing of the package: the CUDA library is never loaded into this process."""
    if rank != 0:
        return
    synth = _synth()
    tau = synth.DEFAULT_TAU
    for _ in range(max(1, args.warmup)):
        cpu_head_rate(REF_SAMPLE_POINTS // 8, tau)
    t_all, pts_all = 0.0, 0.0
    for _ in range(args.steps):
        rate, dt, passes = cpu_head_rate(REF_SAMPLE_POINTS, tau)
        t_all += dt; pts_all += passes * REF_SAMPLE_POINTS
    value = pts_all / t_all
    single = cpu_single_thread_rate(REF_SAMPLE_POINTS // 4, tau)
    cores = blas_threads()
    line = {"impl": "reference", "metric": "points/sec deskew+soft-assign+moment-match (whole evidence chain)", "value": value,
            "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": head_config(tau),
            "scans_per_sec": value / HEAD_POINTS,
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "single_thread_value": single,
                             "sample": f"the first {REF_SAMPLE_POINTS} points of the same {HEAD_POINTS}-point cloud against all {HEAD_BINS} bins per step, "
                                       f"{args.steps} steps; NumPy float64 restatement of the reference operators (JAX unavailable); "
                                       f"os.cpu_count()={os.cpu_count()}, BLAS threads={cores}"},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "loaded_native_libraries": [m for m in ("fl_slam_b200", "fl_slam_b200._abi") if m in sys.modules]}
    emit(line)


_JSON_OUT = None


def claim_stdout():
    """Only the one JSON line may reach stdout (NCCL, for one, prints its version banner there under torchrun): keep a private
    copy of the descriptor for that line and point descriptor 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def bench_stream(args, g, synth, dev, rank, world, local_rank, tau, peak_gbs, peak_src, precision):
    """BASELINE.json configs[1] (side measurement; headline of round 1): the 200-scan stream with map + IW state carried, one
    replica per rank, and the per-scan host calls (pinned host buffers) with their latencies."""
    import torch
    import torch.distributed as dist
    args = argparse.Namespace(**vars(args))
    args.precision = precision
    # ---------------------------------------------------------------- workload: configs[1] scan stream (replica per rank)
    scans, poses = make_stream(N_SCANS, seed0=1234 + 2 + 1000 * rank)
    pl = g.LidarEvidencePipeline(n_bins=N_BINS, tau=tau, precision=args.precision, n_points_cap=CAP,
                                 lidar_origin=scans[0].lidar_origin)
    res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans]
    poses_dev = torch.from_numpy(poses).to(dev)
    xis = np.stack([s.xi_body for s in scans])
    jobs_dev = pl.jobs_to_device([pl.make_job(res[k], xis[k]) for k in range(N_SCANS)])
    ms = g.create_empty_map_stats(N_BINS)
    iw = pl.create_lidar_iw_state(1e-3)        # Kimera value of GC_LIDAR_SIGMA_MEAS (config/gc_unified.yaml:16)
    total_points = sum(s.points.shape[0] for s in scans)
    bpp = res[0].bytes_per_point
    # algorithmic bytes of scan_moments: 5 scalars per SELECTED row + every raw weight when stride > 1
    def _alg_bytes(n):
        stride, n_sel = pl.selected_rows(n, CAP)
        return n_sel * bpp + (n * (bpp // 5) if stride > 1 else 0)
    bytes_per_launch = float(sum(_alg_bytes(s.points.shape[0]) for s in scans))
    max_sel = max(pl.selected_rows(s.points.shape[0], CAP)[1] for s in scans)
    ctas = pl.plan_ctas(N_SCANS, max_sel)

    def stream_pass():
        # scan_moments + reduce_finalize for all 200 scans (independent of the map), then ONE chain launch that walks the
        # scans in order carrying map statistics (forgetting 0.99) and the LiDAR measurement-noise IW state on the device
        # Consecutive batches are pipelined (overlap=True): the moment pass of batch k+1 does not depend on the map, so it runs on
        # 147 SMs while the sequential chain of batch k keeps one SM busy; map and IW state still pass from batch to batch in order.
        pl.run_stream(res, xis, poses_dev, poses_dev, ms, iw_state=iw, use_sigma_meas=True, jobs_dev=jobs_dev, ctas=ctas, overlap=overlap)

    overlap = False
    launches0 = g._abi.launch_count(local_rank)
    stream_pass()                                   # warm-up (allocates work buffers, sets kernel attributes)
    torch.cuda.synchronize()
    launches_per_step = g._abi.launch_count(local_rank) - launches0
    # the same K steps without pipelining (one batch at a time), for the record
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        stream_pass()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        stream_pass()
    ev1.record()
    torch.cuda.synchronize()
    ms_per_step_serial = ev0.elapsed_time(ev1) / args.steps
    overlap = not args.no_overlap
    for _ in range(args.warmup):
        stream_pass()
    pl.join_stream()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        stream_pass()
    pl.join_stream()                                # the current stream waits for the last batch's chain
    ev1.record()
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * total_points / (ms_per_step * 1e-3)

    # ---------------------------------------------------------------- roofline of the dominant kernel (scan_moments alone)
    b = pl._buffers(N_SCANS, ctas)
    def moments_only():
        pl.launch_moments(jobs_dev, None, N_SCANS, ctas, res[0].layout, b["partials"])
    def rest_only():
        pl.launch_finalize(b["partials"], ctas, pl.rec, pl.rec * ctas, jobs_dev, None, N_SCANS, True, b["reduced"], b["stats"])
        pl.launch_chain(b["stats"], N_SCANS, 1, ms, b["map_derived"], poses_dev, poses_dev, b["evidence"], iw, True, 0.99)
    def _time(fn, n=max(args.steps, 100)):
        fn(); torch.cuda.synchronize()
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record(); torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / n
    k1_ms = _time(moments_only)
    rest_ms = _time(rest_only)
    achieved = bytes_per_launch / (k1_ms * 1e-3) / 1e9
    pairs = float(sum(pl.selected_rows(s.points.shape[0], CAP)[1] for s in scans)) * N_BINS
    roofline = {"bound": "hbm", "kernel": "scan_moments_fused_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": None, "peak_source": peak_src, "ms_per_launch": k1_ms,
                "launches_per_step": 1, "algorithmic_bytes_per_launch": bytes_per_launch,
                "share_of_step": k1_ms / ms_per_step, "rest_of_step_ms": rest_ms,
                "pairs_per_sec": pairs / (k1_ms * 1e-3),
                "traffic_source": TRAFFIC_SOURCE,
                "note": "one launch covers all 200 scans (7500 rows x 48 bins each); the kernel is bound by FMA-pipe issue, not by "
                        "HBM: compute_frac = pairs/s over the FMA-issue ceiling of its own instruction mix at the sampled SM "
                        "clock (DESIGN.md section 4).  The other large launch of a step is the sequential measurement-noise core "
                        "(stream_sigma_chain_kernel, one warp walking the 200 scans, ~2.6 us per scan: latency-bound by construction); "
                        "the remaining five launches take ~0.1 ms together (profiles/r1_launches_bench_stream_fp64.csv)"}
    # ---------------------------------------------------------------- e2e: the reference-facing host call, pinned buffers
    n_e2e = min(N_SCANS, 50)
    pinned = []
    for s in scans[:n_e2e]:
        p, t_, w_ = (torch.from_numpy(a).pin_memory() for a in (s.points, s.timestamps, s.weights))
        pinned.append((p.numpy(), t_.numpy(), w_.numpy(), p, t_, w_))
        pl.register_pinned(*pinned[-1][:3])     # a ring of page-locked receive buffers, registered once

    lat_blocking = []          # per-scan latency of the blocking call (SURVEY.md 8d asks for p50 / p99 on configs[1])

    def e2e_pass():
        m = {k: np.zeros(v) for k, v in (("S_dir", (N_BINS, 3)), ("S_dir_scatter", (N_BINS, 3, 3)), ("N_dir", (N_BINS,)),
                                         ("N_pos", (N_BINS,)), ("sum_p", (N_BINS, 3)), ("sum_ppT", (N_BINS, 3, 3)))}
        iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
        for k in range(n_e2e):
            t_a = time.perf_counter()
            pl.process_scan_host(pinned[k][0], pinned[k][1], pinned[k][2], scans[k].scan_start_time, scans[k].scan_end_time,
                                 xis[k][None], poses[k][None], m, pose_update=poses[k], iw_state_host=iw_h, use_sigma_meas=True)
            lat_blocking.append(time.perf_counter() - t_a)
        return m
    def e2e_pass_pipelined():
        """Same work per scan, two-stage API: the upload + moment pass of scans k+1 and k+2 (which need only the scan and the IMU
        twist) are in flight while scan k's evidence step (which needs the pose prediction and the map) runs."""
        m = {k: np.zeros(v) for k, v in (("S_dir", (N_BINS, 3)), ("S_dir_scatter", (N_BINS, 3, 3)), ("N_dir", (N_BINS,)),
                                         ("N_pos", (N_BINS,)), ("sum_p", (N_BINS, 3)), ("sum_ppT", (N_BINS, 3, 3)))}
        iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
        sub = lambda k: pl.submit_scan_host(k % 4, pinned[k][0], pinned[k][1], pinned[k][2], scans[k].scan_start_time,
                                            scans[k].scan_end_time, xis[k][None])
        sub(0); sub(1)
        for k in range(n_e2e):
            pl.evidence_begin_host(k % 4, poses[k][None], m, pose_update=poses[k], iw_state_host=iw_h, use_sigma_meas=True)
            if k + 2 < n_e2e:
                sub(k + 2)                      # an upload + moment pass two scans ahead goes out while the chain step runs
            pl.evidence_end_host()
        return m

    # the same stream from the MESSAGES: 22 bytes per point of raw PointCloud2 + the 512-sample IMU buffer per scan cross PCIe;
    # parse, range weights, base-frame transform, scan time bounds, IMU preintegration -> twist, resample, moments on the device
    R_bl, t_bl = synth.rotation_from_rotvec(np.array([0.02, -0.01, 0.3])), scans[0].lidar_origin.copy()
    msgs = []
    for k, s in enumerate(scans[:n_e2e]):
        m = synth.make_pointcloud2(s, R_bl, t_bl, time_field="time_f32", time_base=s.scan_start_time - 0.1 * k - 5.0)
        buf = torch.from_numpy(m["data"]).pin_memory()
        imu = synth.make_imu_buffer(5.0 + 0.1 * k, 5.1 + 0.1 * k, 512, seed=k)
        imu.update(sigma=0.01, rotvec_start_WB=np.zeros((1, 3)), gyro_bias=np.zeros((1, 3)), accel_bias=np.zeros((1, 3)),
                   gravity_W=np.array([0.0, 0.0, -9.81]))
        msgs.append((buf.numpy(), m, imu, buf))

    def e2e_pass_messages():
        m = {k: np.zeros(v) for k, v in (("S_dir", (N_BINS, 3)), ("S_dir_scatter", (N_BINS, 3, 3)), ("N_dir", (N_BINS,)),
                                         ("N_pos", (N_BINS,)), ("sum_p", (N_BINS, 3)), ("sum_ppT", (N_BINS, 3, 3)))}
        iw_h = np.concatenate([[4.5], (1e-3 * 0.5 * np.eye(3)).reshape(-1)])
        sub = lambda k: pl.submit_scan_msg_host(k % 4, data=msgs[k][0], fields=msgs[k][1]["fields"], point_step=msgs[k][1]["point_step"],
                                                n_points=msgs[k][1]["n_points"], header_stamp_sec=msgs[k][1]["header_stamp_sec"],
                                                R_base_lidar=R_bl, t_base_lidar=t_bl, imu=msgs[k][2])
        sub(0); sub(1)
        for k in range(n_e2e):
            pl.evidence_begin_host(k % 4, poses[k][None], m, pose_update=poses[k], iw_state_host=iw_h, use_sigma_meas=True)
            if k + 2 < n_e2e:
                sub(k + 2)
            pl.evidence_end_host()
        return m

    def _time_e2e(fn):
        fn()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_rep):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt
    n_rep = max(1, args.steps // 2)
    e2e_blocking_s = _time_e2e(e2e_pass)
    lat = np.sort(np.array(lat_blocking[n_e2e:])) * 1e6          # the first pass is the warm-up
    e2e_s = _time_e2e(e2e_pass_pipelined)
    e2e_msg_s = _time_e2e(e2e_pass_messages)
    e2e_points = sum(s.points.shape[0] for s in scans[:n_e2e])
    rec_bytes = (38 * N_BINS + 24 + 640 + 26 * N_BINS + 10) * 8
    e2e = {"value": world * n_rep * e2e_points / e2e_s, "unit": "points/s",
           "h2d_bytes_per_step": int(sum(_alg_bytes(s.points.shape[0]) for s in scans[:n_e2e]) + n_e2e * (48 + 48 + 120)),
           "d2h_bytes_per_step": int(n_e2e * rec_bytes), "scans_per_sec": world * n_rep * n_e2e / e2e_s,
           "us_per_scan": 1e6 * e2e_s / (n_rep * n_e2e),
           "blocking_call": {"value": world * n_rep * e2e_points / e2e_blocking_s, "unit": "points/s",
                             "us_per_scan": 1e6 * e2e_blocking_s / (n_rep * n_e2e),
                             "latency_us": {"p50": float(lat[len(lat) // 2]), "p99": float(lat[min(len(lat) - 1, int(0.99 * len(lat)))]),
                                            "max": float(lat[-1]), "n": int(len(lat))},
                             "note": "one gcl_lidar_scan_evidence_host call per scan, nothing overlapped"},
           "from_message": {"value": world * n_rep * e2e_points / e2e_msg_s, "unit": "points/s", "us_per_scan": 1e6 * e2e_msg_s / (n_rep * n_e2e),
                            "h2d_bytes_per_step": int(e2e_points * msgs[0][1]["point_step"] + n_e2e * (512 * 7 * 8 + N_BINS * 26 * 8 + N_BINS * 24 + 400)),
                            "note": "gcl_scan_submit_msg_host: the raw PointCloud2 bytes (22 B/point, Kimera VLP-16 layout) and the 512-sample IMU "
                                    "buffer are the inputs; parse, range weights, base-frame transform, scan time bounds, IMU preintegration -> "
                                    "deskew twist, resample, moments and finalize run on the device (the reference does the parse in NumPy and the "
                                    "preintegration as a 512-step lax.scan before the operators timed by the other figures)"},
           "note": f"{n_e2e} scans per pass through the two-stage host API with pinned host arrays: gcl_scan_submit_host (H2D of "
                   "the scan, scan_moments, reduce_finalize; asynchronous) for scan k+2 is issued between "
                   "gcl_scan_evidence_begin_host (H2D of pose/map/IW, chain step incl. IW + map update, one D2H of "
                   "map/IW/stats/evidence enqueued) and gcl_scan_evidence_end_host (sync, results out) of scan k; every scan's inputs cross PCIe and every scan's results are read back inside the timed region. "
                   "Stage 1 is issued by the handle's submit thread and reads the page-locked scan arrays in place (zero-copy: scan_moments pulls the "
                   "selected rows and every raw weight across PCIe once, which is what h2d_bytes_per_step counts); stage 2 is ONE launch: poses as "
                   "kernel arguments, map / IW fetched only when they differ from what the device already holds (here they never do: the arrays "
                   "handed back are the ones the previous step wrote), map / IW / stats / evidence pushed back by the kernel, then a ticket the host polls"}


    roofline["traffic"] = TRAFFIC_BYTES_PER_LAUNCH.get(precision)
    ceiling = pair_ceiling(precision, 1965.0)
    roofline["compute_ceiling_pairs_per_sec_at_1965MHz"] = ceiling
    roofline["compute_frac_at_1965MHz"] = roofline["pairs_per_sec"] / ceiling
    out = {"workload": f"configs[1]: scan stream, {N_SCANS} scans x ~{PTS_PER_SCAN} points, cap {CAP} (stride 4, 7500 rows), {N_BINS} bins, "
                       "1 hypothesis, map statistics carried (forgetting 0.99), LiDAR measurement-noise IW block adapted every scan",
           "precision": precision, "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas (a stream does not shard)",
           "points_per_sec": value, "scans_per_sec": world * N_SCANS / (ms_per_step * 1e-3), "ms_per_step": ms_per_step,
           "ms_per_step_one_batch_at_a_time": ms_per_step_serial, "launches_per_step": int(launches_per_step),
           "processed_rows_per_sec": value * 7500.0 / PTS_PER_SCAN, "roofline": roofline, "e2e": e2e}
    del res, pinned, msgs
    return out


def main():
    global HEAD_POINTS, HEAD_BINS
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32c"])
    ap.add_argument("--no-extra", action="store_true", help="skip the fp32c / configs[1] / configs[2] side measurements")
    ap.add_argument("--no-overlap", action="store_true", help="stream side measurement: one batch at a time")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: fused NVLink exchange kernel, or the NCCL all-gather route")
    ap.add_argument("--scans-per-step", type=int, default=SCANS_PER_STEP)
    ap.add_argument("--head-points", type=int, default=HEAD_POINTS, help="points per scan of the sharded cloud (default = configs[3]; "
                    "16777216 with --head-bins 1024 = configs[4], the extreme sweep)")
    ap.add_argument("--head-bins", type=int, default=HEAD_BINS)
    args = ap.parse_args()
    custom_head = (args.head_points, args.head_bins) != (HEAD_POINTS, HEAD_BINS)
    HEAD_POINTS, HEAD_BINS = args.head_points, args.head_bins
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import fl_slam_b200 as g
    from fl_slam_b200 import synth
    from fl_slam_b200 import sharding

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host = None
    if world > 1:
        # one process per GPU: keep this rank's threads and page-locked buffers on the GPU's own NUMA node (N = 1 is left alone: the CPU
        # baseline runs in that process and must keep every core)
        from fl_slam_b200 import hostmem
        host = hostmem.bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=dev)
    tau = synth.DEFAULT_TAU
    peak_gbs, peak_src = load_peaks()

    # ---------------------------------------------------------------- headline: configs[3], points sharded over the ranks
    sampler = ClockSampler(local_rank)
    sampler.start()
    head = sharding.bench_points_sharded(g, n_points=HEAD_POINTS, n_bins=HEAD_BINS, tau=tau, steps=args.steps, warmup=args.warmup, rank=rank,
                                         world=world, peak_gbs=peak_gbs, precision=args.precision, scans_per_step=args.scans_per_step,
                                         exchange=args.exchange, with_e2e=not args.no_e2e)
    clocks = sampler.stop()
    roofline = head.pop("roofline")
    roofline["peak_source"] = peak_src
    roofline["traffic"] = HEAD_TRAFFIC_BYTES_PER_LAUNCH.get(args.precision) if world == 1 and not custom_head else None
    roofline["traffic_source"] = HEAD_TRAFFIC_SOURCE
    ceiling = pair_ceiling(args.precision, clocks.get("sm_mhz") or 1965.0)
    roofline["pairs_per_sec"] = head["pairs_per_sec"] / world       # per GPU, like achieved
    roofline["compute_ceiling_pairs_per_sec"] = ceiling
    roofline["compute_frac"] = roofline["pairs_per_sec"] / ceiling
    roofline["note"] = ("scan_moments is bound by FMA-pipe issue, not by HBM (DESIGN.md section 4): compute_frac = pairs/s over the FMA-issue "
                        "ceiling of its own instruction mix at the sampled SM clock; frac is the HBM fraction the north star asks for")
    e2e = head.pop("e2e", None)
    parity = head.pop("parity")

    # ---------------------------------------------------------------- side measurements
    extra = {}
    if not args.no_extra:
        other = "fp32c" if args.precision == "fp64" else "fp64"
        o = sharding.bench_points_sharded(g, n_points=HEAD_POINTS, n_bins=HEAD_BINS, tau=tau, steps=max(3, args.steps // 2), warmup=args.warmup,
                                          rank=rank, world=world, peak_gbs=peak_gbs, precision=other, scans_per_step=args.scans_per_step,
                                          exchange=args.exchange, with_e2e=not args.no_e2e)
        o["roofline"]["compute_frac"] = (o["pairs_per_sec"] / world) / pair_ceiling(other, clocks.get("sm_mhz") or 1965.0)
        o["roofline"]["compute_frac_note"] = ("pairs/s over the FMA-issue ceiling of the CUDA-core instruction mix (24 lane operations per pair); the "
                                              "tensor-core kernel moves 19 of them to tcgen05 and is bound by the issue slots of what remains "
                                              "(DESIGN.md, K1-TC)") if other == "fp32c" else "pairs/s over the FP64 FMA-issue ceiling (31 lane operations per pair)"
        extra[f"cfg4_1Mx256_{other}"] = o
        if world > 1 and args.exchange == "peer":
            n = sharding.bench_points_sharded(g, n_points=HEAD_POINTS, n_bins=HEAD_BINS, tau=tau, steps=max(3, args.steps // 2), warmup=args.warmup,
                                              rank=rank, world=world, peak_gbs=peak_gbs, precision=args.precision,
                                              scans_per_step=args.scans_per_step, exchange="nccl", with_e2e=False)
            extra["cfg4_1Mx256_nccl_all_gather_route"] = {k: n[k] for k in ("ms_per_scan", "scan_moments_ms", "rest_of_scan_ms", "points_per_sec", "parity")}
        extra["cfg3_64hyp"] = sharding.bench_hypothesis_sharded(g, n_points=30000, cap=32768, n_bins=48, n_hyp=64, tau=tau,
                                                               steps=max(3, args.steps // 2), warmup=args.warmup, rank=rank, world=world)
        extra["cfg3_64hyp_fp32c"] = sharding.bench_hypothesis_sharded(g, n_points=30000, cap=32768, n_bins=48, n_hyp=64, tau=tau,
                                                                     steps=max(3, args.steps // 2), warmup=args.warmup, rank=rank, world=world,
                                                                     precision="fp32c")
        extra["cfg2_stream"] = bench_stream(args, g, synth, dev, rank, world, local_rank, tau, peak_gbs, peak_src, "fp64")

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only) + parity of the GPU path on its sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = REF_SAMPLE_POINTS
        rate, dt, passes, (sc_s, raw_s, fin_s, ms_s, mf_s, pt_s) = cpu_head_rate(n_s, tau, min_seconds=12.0, return_result=True)
        single = cpu_single_thread_rate(n_s // 4, tau)
        pl_s = g.LidarEvidencePipeline(n_bins=HEAD_BINS, tau=tau, precision=args.precision, n_points_cap=n_s, lidar_origin=sc_s.lidar_origin)
        ev_s = pl_s.process_scan_host(sc_s.points[:n_s], sc_s.timestamps[:n_s], sc_s.weights[:n_s], sc_s.scan_start_time, sc_s.scan_end_time,
                                      sc_s.xi_body[None], np.zeros((1, 6)), ms_s)
        sb = ev_s.scan_bins(0)
        def _r(a, b):
            a, b = np.asarray(a.cpu().numpy() if hasattr(a, "cpu") else a), np.asarray(b)
            return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
        chk = {"N": _r(sb.N, fin_s["N"]), "s_dir": _r(sb.s_dir, fin_s["s_dir"]), "S_dir_scatter": _r(sb.S_dir_scatter, fin_s["S_dir_scatter"]),
               "p_bar": _r(sb.p_bar, fin_s["p_bar"]), "Sigma_p": _r(sb.Sigma_p, fin_s["Sigma_p"]),
               "R_mf_rad": float(np.linalg.norm(ev_s.evidence[0, :9].cpu().numpy().reshape(3, 3) - mf_s["R_mf"]) / np.sqrt(2.0)),
               "t_wls": _r(ev_s.evidence[0, 71:74], pt_s["t_wls"])}
        cpu = {"value": rate, "unit": "points/s", "cores": blas_threads(), "kind": "port", "single_thread_value": single,
               "sample": f"the first {n_s} points of the same {HEAD_POINTS}-point cloud against all {HEAD_BINS} bins, walked {passes} times in {dt:.1f} s; "
                         f"NumPy float64 restatement of the reference operators (JAX unavailable); os.cpu_count()={os.cpu_count()}",
               "gpu_vs_oracle_on_sample_max_rel": chk}
        parity["vs_oracle_on_cpu_sample"] = chk

    if rank == 0:
        cfg = head_config(tau, args.precision)
        line = {"metric": "points/sec deskew+soft-assign+moment-match (whole evidence chain)", "value": head["points_per_sec"], "unit": "points/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64" if args.precision == "fp64" else "f32 (fp64 carry)",
                "data": "synthetic", "config": cfg,
                "details": {"layout": "ref_f64 (40 B/point)" if args.precision == "fp64" else "packed_f32 (20 B/point)",
                            "parallelism": "1 GPU" if world == 1 else f"points sharded over {world} GPUs, one fused NVLink exchange of the "
                                           f"{(19 * HEAD_BINS + 8) * 8} B per-bin record per scan ({head['exchange']})",
                            "step": f"{head['scans_per_step']} scans", "l2": f"{head['input_copies_rotated']} copies of the input rotated "
                                    "(> 2 x the 126 MB L2 per rank): no scan re-reads an L2-resident cloud",
                            "timed_region_ms": head["ms_per_step"] * args.steps, **{k: head[k] for k in (
                                "ms_per_scan", "ms_per_scan_median", "ms_per_scan_best", "ms_per_scan_one_at_a_time", "pipelining", "scan_moments_ms",
                                "rest_of_scan_ms", "launches_per_scan",
                                "ctas_per_rank", "rows_per_rank")}},
                "scans_per_sec": head["scans_per_sec"], "pairs_per_sec": head["pairs_per_sec"],
                "processed_rows_per_sec": head["points_per_sec"], "gpu_launches": int(head["launches_per_step"] * args.steps),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "clocks": clocks, "host_placement": host, "extra": extra}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
