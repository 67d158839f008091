"""The sharded CUDA paths on real GPUs (needs >= 2; skipped on a one-GPU box): one process per GPU, NCCL only for the rendezvous and
the checks, the exchange itself through the fused NVLink kernel (csrc/gcl_exchange.cu).

  * points sharded (BASELINE.json configs[3] shape, reduced): PointShardedEvidence.run on every rank == the same job unsharded
    (every raw sum is linear in the points, archive/legacy_operators/binning.py:166-173: only the grouping of the fp64 additions
    differs -> <= 1e-12), bit-identical on every rank (docs/GC_SLAM.md:1150 determinism), with stride > 1 (mass rescale after the
    exchange, point_budget.py:80-85), several hypotheses, the round-1 NCCL route as a cross-check, consecutive scans in flight through the
    scan pipe from resident shards and from page-locked host rows;
  * hypotheses sharded (configs[2] shape): rank r's slice of the batch == the full batch's rows, gathered through the exchange.
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rel(a, b):
    d = float((a - b).abs().max())
    m = float(b.abs().max())
    return d / (m if m > 0 else 1.0)


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GCL_EXCHANGE_TIMEOUT_S="20")
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import fl_slam_b200 as g
    from fl_slam_b200 import sharding as S, synth
    cases = [dict(n=200_003, cap=200_003, B=256, H=1, dense=True, precision="fp64"),     # stride 1, ragged split
             dict(n=30_011, cap=8192, B=48, H=1, dense=False, precision="fp64"),         # stride 4: raw-range ownership + mass rescale
             dict(n=50_000, cap=65536, B=48, H=3, dense=False, precision="fp64"),        # several jobs through one exchange
             dict(n=200_003, cap=200_003, B=256, H=1, dense=True, precision="fp32c")]
    for c in cases:
        sc = synth.make_scan(c["n"], seed=4321, dense=c["dense"])
        pl = g.LidarEvidencePipeline(n_bins=c["B"], tau=0.1, precision=c["precision"], n_points_cap=c["cap"], lidar_origin=sc.lidar_origin)
        rng = np.random.default_rng(7)
        xis = sc.xi_body[None] * (1 + 0.02 * rng.normal(size=(c["H"], 6)))
        poses = torch.from_numpy(0.02 * rng.normal(size=(c["H"], 6))).cuda()
        full = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
        ms0 = g.create_empty_map_stats(c["B"])
        ref0 = pl.run_resident(full, xis[:1], poses[:1], ms0)
        sa = {k: v.detach().cpu().numpy() for k, v in S._stats_arrays(pl, ref0.stats[0]).items()}
        raw = {"N": sa["N"], "s_dir": sa["s_dir"].reshape(-1, 3), "S_dir_scatter": sa["S_dir_scatter"].reshape(-1, 3, 3),
               "sum_p": sa["sum_p"].reshape(-1, 3), "sum_ppT": sa["sum_ppT"].reshape(-1, 3, 3)}
        ms = g.MapBinStats.from_arrays(synth.map_stats_from_scan_sums(raw, *synth.known_map_transform(1), seed=1))
        ref = pl.run_resident(full, xis, poses, ms)
        ref_stats, ref_evid = ref.stats.clone(), ref.evidence.clone()
        tol = 1e-12 if c["precision"] == "fp64" else 2e-6
        for route in ("peer", "nccl"):
            sh = S.PointShardedEvidence(pl, rank, world, exchange=route, max_jobs=c["H"])
            shard, n_raw = sh.upload_shard(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
            for rep in range(3):                       # several steps through the same exchange (parity of the buffers alternates)
                ev = sh.run(shard, n_raw, xis, poses, ms)
            torch.cuda.synchronize()
            assert sh.px is None or sh.px.error_flag() == 0
            for h in range(c["H"]):
                a, b = S._stats_arrays(pl, ev.stats[h]), S._stats_arrays(pl, ref_stats[h])
                for k in b:
                    lim = tol if k not in ("Sigma_p", "kappa") else (1e-9 if c["precision"] == "fp64" else 1e-3)
                    assert _rel(a[k], b[k]) < lim, (c, route, h, k, _rel(a[k], b[k]))
                sc_a, sc_b = ev.stats[h, 38 * c["B"]:], ref_stats[h, 38 * c["B"]:]
                assert _rel(sc_a[:16], sc_b[:16]) < (1e-11 if c["precision"] == "fp64" else 1e-5), (c, route, h)
                assert float((ev.evidence[h, 0:9] - ref_evid[h, 0:9]).norm()) < (1e-9 if c["precision"] == "fp64" else 1e-4), (c, route, h)
            assert S._bits_identical_across_ranks(torch.cat([ev.stats.reshape(-1), ev.evidence.reshape(-1)]), world), (c, route)
            if route == "peer":
                # consecutive scans in flight (gcl_scan_pipe: two tail streams on the two exchange lanes, four buffer sets): every scan
                # equals the same scan run alone with the same grid, bit for bit, twists differing from scan to scan
                sms = int(g._abi.lib.gcl_device_sm_count(pl.handle))
                _, _, (b0, e0), _ = sh.plan(n_raw)
                ctas = max(1, min(pl.plan_ctas(c["H"], max(1, e0 - b0)), (sms - 8) // c["H"]))
                variants = [xis * (1 + 0.03 * v) for v in range(2)]
                want = []
                for v in range(2):
                    e_ = sh.run(shard, n_raw, variants[v], poses, ms, ctas=ctas)
                    want.append((e_.stats.clone(), e_.evidence.clone()))
                order = [0, 1, 1, 0, 0, 1, 0, 1]
                got = []
                for k, v in enumerate(order):
                    got.append(sh.run_overlapped(shard, n_raw, variants[v], poses, ms, reserve_sms=8, n_tails=2))
                    if k % 4 == 3:
                        sh.join_overlapped()
                        torch.cuda.synchronize()
                        for kk in range(k - 3, k + 1):
                            assert torch.equal(got[kk].stats, want[order[kk]][0]), (c, "pipelined stats", kk)
                            assert torch.equal(got[kk].evidence, want[order[kk]][1]), (c, "pipelined evidence", kk)
                if c["H"] == 1:
                    # the same scans from page-locked HOST rows through the pipe (run_host_pipelined: uploads ordered by
                    # gcl_scan_pipe_wait_set, D2H of the records at the end of each tail): six scans over the four sets, every read-back
                    # record bit-equal to the resident pipelined run, and the same bits on every rank
                    st = sh.host_stage_pipelined(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time, reserve_sms=8,
                                                 n_tails=2)
                    sets = [sh.run_host_pipelined(st, variants[0], poses, ms)[1] for _ in range(6)]
                    sh.join_overlapped()
                    torch.cuda.synchronize()
                    for i_ in set(sets[2:]):
                        assert torch.equal(st["out"][i_][0], want[0][0].cpu()), (c, "host pipelined stats", i_)
                        assert torch.equal(st["out"][i_][1], want[0][1].cpu()), (c, "host pipelined evidence", i_)
                    assert S._bits_identical_across_ranks(torch.cat([st["out"][sets[-1]][0].reshape(-1), st["out"][sets[-1]][1].reshape(-1)]).cuda(), world)
                    sh.end_host_pipelined(st)
                sh.end_overlapped()
                assert sh.px.error_flag() == 0
            if sh.px is not None:
                sh.px.close()
        del full
    # ---- hypotheses sharded + result gather through the exchange -------------------------------------------------------------
    n, cap, B, Hn = 30_000, 32768, 48, 8 * world
    sc = synth.make_scan(n, seed=99)
    pl = g.LidarEvidencePipeline(n_bins=B, tau=0.1, precision="fp64", n_points_cap=cap, lidar_origin=sc.lidar_origin)
    rng = np.random.default_rng(3)
    xis = sc.xi_body[None] * (1 + 0.02 * rng.normal(size=(Hn, 6)))
    poses = torch.from_numpy(0.02 * rng.normal(size=(Hn, 6))).cuda()
    scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
    ms = g.create_empty_map_stats(B)
    allev = pl.run_resident(scan, xis, poses, ms).evidence.clone()
    lo, hi = S.shard_rows(Hn, rank, world)
    mine = pl.run_resident(scan, xis[lo:hi], poses[lo:hi], ms).evidence
    px = S.PeerExchange(pl, rank, world, (hi - lo) * 640 * 8)
    gathered = torch.empty((world, hi - lo, 640), dtype=torch.float64, device="cuda")
    for rep in range(3):
        px.all_gather(mine, gathered)
    torch.cuda.synchronize()
    assert px.error_flag() == 0
    # (the number of CTAs per job depends on the batch size, so a slice of the batch regroups the fp64 partial sums: not bit-equal)
    assert _rel(gathered.reshape(Hn, 640), allev) < 1e-9, "gathered hypothesis records differ from the unsharded batch"
    assert torch.equal(gathered[rank], mine)
    assert S._bits_identical_across_ranks(gathered.reshape(-1), world)
    # the same through the pipe (hypothesis-batch mode: finalise -> evidence -> gather on the tail streams, lanes alternating)
    jobs_dev = pl.jobs_to_device([pl.make_job(scan, xis[h]) for h in range(lo, hi)])
    cj = pl.plan_ctas(hi - lo, 30_000)
    pipe = S.ScanPipe(pl, px, world, "gather", n_sets=4, n_tails=2)
    outs = [pipe.submit(jobs_dev, None, hi - lo, cj, scan.layout, ms, poses[lo:hi]) for _ in range(6)]
    pipe.join()
    torch.cuda.synchronize()
    assert px.error_flag() == 0
    for buf, gat in outs[2:]:
        assert _rel(buf["evidence"], mine) < 1e-12
        assert torch.equal(gat[rank], buf["evidence"]) and _rel(gat.reshape(Hn, 640), allev) < 1e-9
        assert S._bits_identical_across_ranks(gat.reshape(-1), world)
    pipe.close()
    px.close()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_cuda_paths_match_unsharded_and_are_bit_identical_across_ranks():
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    world = min(n_gpu, 4)
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port()), nprocs=world, join=True)
