#!/usr/bin/env python
"""What bounds the pipelined point-sharded scan when the per-rank shard is small (the 8-GPU shard of configs[3] on ONE GPU, world = 1):
host enqueue time per scan, the moment pass alone, the exchange + finalise + evidence tail alone, and the pipelined scan.

    python tools/tail_probe.py [n_rows_per_rank] [n_bins] [reserve:tails,reserve:tails,...]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fl_slam_b200 as g  # noqa: E402
from fl_slam_b200.sharding import PointShardedEvidence  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
CASES = tuple(tuple(int(v) for v in c.split(":")) for c in sys.argv[3].split(",")) if len(sys.argv) > 3 else ((6, 1), (6, 2))
out = {}
for precision in ("fp32c", "fp64"):
    sc = g.synth.make_scan(n, seed=1238, dense=True)
    pl = g.LidarEvidencePipeline(n_bins=B, tau=g.synth.DEFAULT_TAU, precision=precision, n_points_cap=n, lidar_origin=sc.lidar_origin)
    sh = PointShardedEvidence(pl, 0, 1, exchange="peer")
    shards = [sh.upload_shard(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)[0] for _ in range(8)]
    ms = g.create_empty_map_stats(B)
    poses = torch.zeros(1, 6, dtype=torch.float64, device="cuda")
    xi = sc.xi_body[None]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    jobs_pre = [sh.make_sharded_job(s_, n, sc.xi_body) for s_ in shards]
    for reserve, tails in CASES:
        for k in range(16):
            sh.run_overlapped(shards[k % 8], n, xi, poses, ms, reserve_sms=reserve, n_tails=tails, job=jobs_pre[k % 8])
        sh.join_overlapped(); torch.cuda.synchronize()
        N = 256
        e0.record()
        t0 = time.perf_counter()
        for k in range(N):
            sh.run_overlapped(shards[k % 8], n, xi, poses, ms, reserve_sms=reserve, n_tails=tails, job=jobs_pre[k % 8])
        sh.join_overlapped()
        t1 = time.perf_counter()
        e1.record(); torch.cuda.synchronize()
        res[f"pipelined_reserve{reserve}_tails{tails}"] = {"host_enqueue_us_per_scan": (t1 - t0) / N * 1e6, "gpu_us_per_scan": e0.elapsed_time(e1) / N * 1e3}
        sh.end_overlapped()
    # one at a time, with events around the parts
    ctas = pl.plan_ctas(1, n)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tk, tt = [], []
    for k in range(40):
        ev = sh.run(shards[k % 8], n, xi, poses, ms, ctas=ctas)
    torch.cuda.synchronize()
    from fl_slam_b200 import _abi, operators as ops
    import ctypes as C
    for k in range(40):
        scan = shards[k % 8]
        buf = pl._buffers(1, ctas)
        job = pl.make_job(scan, xi[0], n_raw_total=n, sel_range=(0, n))
        evs[0].record()
        pl.launch_moments(None, job, 1, ctas, scan.layout, buf["partials"])
        evs[1].record()
        sh.px._check(_abi.lib.gcl_sharded_reduce_finalize(
            pl.handle, sh.px.x, ops._stream(), ops._ptr(buf["partials"]), ctas, pl.rec, pl.rec * ctas, None, C.byref(job), 1, pl.n_bins,
            ops._ptr(pl.bin_dirs), pl.tau, pl._origin_c, pl.eps_psd, pl.eps_mass, ops._ptr(buf["reduced"]), ops._ptr(buf["stats"])))
        evs[2].record()
        pl.launch_chain(buf["stats"], 1, 1, ms, buf["map_derived"], poses, None, buf["evidence"])
        e1.record()
        torch.cuda.synchronize()
        tk.append((evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(e1)))
    a = np.median(np.asarray(tk), axis=0) * 1e3
    res["one_at_a_time_us"] = {"scan_moments": a[0], "exchange_finalize": a[1], "chain": a[2], "ctas": ctas}
    out[precision] = res
    sh.px.close()
# the tail of a hypothesis batch as one rank of 8 sees it: 8 hypotheses x 17 CTAs, 48 bins
for precision in ("fp64",):
    sc = g.synth.make_scan(30000, seed=1237)
    pl = g.LidarEvidencePipeline(n_bins=48, tau=g.synth.DEFAULT_TAU, precision=precision, n_points_cap=32768, lidar_origin=sc.lidar_origin)
    scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
    Hn, cj = 8, 17
    rng = np.random.default_rng(1)
    xis = sc.xi_body[None] * (1 + 0.02 * rng.normal(size=(Hn, 6)))
    poses = torch.from_numpy(0.02 * rng.normal(size=(Hn, 6))).cuda()
    jobs = pl.jobs_to_device([pl.make_job(scan, xis[h]) for h in range(Hn)])
    ms = g.create_empty_map_stats(48)
    b = pl._buffers(Hn, cj)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tk = []
    for k in range(60):
        evs[0].record()
        pl.launch_moments(jobs, None, Hn, cj, scan.layout, b["partials"])
        evs[1].record()
        pl.launch_finalize(b["partials"], cj, pl.rec, pl.rec * cj, jobs, None, Hn, True, b["reduced"], b["stats"])
        evs[2].record()
        pl.launch_chain(b["stats"], 1, Hn, ms, b["map_derived"], poses, None, b["evidence"])
        evs[3].record()
        torch.cuda.synchronize()
        if k >= 20:
            tk.append((evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])))
    a = np.median(np.asarray(tk), axis=0) * 1e3
    out["hyp8x17_48bins_" + precision] = {"scan_moments": a[0], "reduce_finalize": a[1], "chain": a[2]}
print(json.dumps(out, indent=1))
