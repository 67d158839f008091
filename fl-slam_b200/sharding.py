"""Multi-GPU execution of the evidence path, one process per GPU (torch.distributed, NCCL over NVLink).

What shards:
  * by points   -- every raw sum and certificate partial is linear in the points (max_resp is a max), so each rank runs
                   the fused kernel on a contiguous range of the SELECTED-row index space and the ranks exchange ONE
                   small record per job ([19 x B moments | 8 scalars], 39 KB at B = 256).  The exchange is an
                   all-gather followed by a fixed-rank-order sum on every rank (deterministic and bit-identical on all
                   ranks); the global mass rescale of PointBudgetResample is applied after it.
  * by hypothesis -- independent jobs; results are all-gathered at the end.
  * a scan stream does not shard (scan k+1 needs the map after scan k): replicas only.
The exchange helper works on CPU tensors with the gloo backend, which is how the N > 1 logic is tested without GPUs.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

NM, NCERT = 19, 8
CS_MAX_RESP = 5


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the selected-row index space owned by `rank`."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def raw_range(sel_begin: int, sel_end: int, stride: int, n_sel: int, n_raw: int) -> Tuple[int, int]:
    """Raw rows a rank must hold: its selected rows and, for mass preservation (sum of ALL raw weights,
    point_budget.py:69), the skipped rows up to the next rank's first selected row."""
    return sel_begin * stride, (n_raw if sel_end >= n_sel else sel_end * stride)


def record_doubles(n_bins: int) -> int:
    return NM * n_bins + NCERT


def exchange_records(reduced: torch.Tensor, group=None) -> torch.Tensor:
    """(n_jobs, rec) on every rank -> (world, n_jobs, rec) on every rank, rank-major.  One collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return reduced.unsqueeze(0)
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(reduced.shape), dtype=reduced.dtype, device=reduced.device)
    if reduced.is_cuda:
        dist.all_gather_into_tensor(out, reduced.contiguous(), group=group)
    else:
        parts = [torch.empty_like(reduced) for _ in range(world)]
        dist.all_gather(parts, reduced.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def combine_records_host(gathered: np.ndarray, n_bins: int) -> np.ndarray:
    """Host restatement of what gcl_reduce_finalize step 1 does with gathered records: rank-order sum, max for the
    max-responsibility slot.  Used by the CPU (gloo) tests of the exchange logic."""
    out = gathered[0].copy()
    mslot = NM * n_bins + CS_MAX_RESP
    for r in range(1, gathered.shape[0]):
        mx = np.maximum(out[..., mslot], gathered[r][..., mslot])
        out = out + gathered[r]
        out[..., mslot] = mx
    return out


class PointShardedEvidence:
    """configs[3]/[4]: one scan, points sharded over the ranks of the default process group."""

    def __init__(self, pipeline, rank: int, world: int):
        self.pl, self.rank, self.world = pipeline, rank, world

    def plan(self, n_raw: int):
        stride, n_sel = self.pl.selected_rows(n_raw, self.pl.n_points_cap)
        b, e = shard_rows(n_sel, self.rank, self.world)
        lo, hi = raw_range(b, e, stride, n_sel, n_raw)
        return stride, n_sel, (b, e), (lo, hi)

    def upload_shard(self, points, timestamps, weights, scan_start_time, scan_end_time):
        """Keeps only this rank's raw rows in HBM."""
        n_raw = int(np.asarray(points).shape[0])
        _, _, _, (lo, hi) = self.plan(n_raw)
        scan = self.pl.upload_scan(points[lo:hi], timestamps[lo:hi], weights[lo:hi], scan_start_time, scan_end_time, row_offset=lo)
        return scan, n_raw

    def run(self, scan, n_raw_total: int, xi_bodies, poses_dev, map_stats, ctas: Optional[int] = None):
        from .pipeline import LidarScanEvidence
        pl = self.pl
        xi = np.asarray(xi_bodies, dtype=np.float64).reshape(-1, 6)
        H = xi.shape[0]
        _, n_sel, (b, e), _ = self.plan(n_raw_total)
        ctas = pl.plan_ctas(H, max(1, e - b)) if ctas is None else ctas
        buf = pl._buffers(H, ctas)
        jobs = [pl.make_job(scan, xi[h], n_raw_total=n_raw_total, sel_range=(b, e)) for h in range(H)]
        job, jobs_dev = (jobs[0], None) if H == 1 else (None, pl.jobs_to_device(jobs))
        pl.launch_moments(jobs_dev, job, H, ctas, scan.layout, buf["partials"])
        pl.launch_finalize(buf["partials"], ctas, pl.rec, pl.rec * ctas, jobs_dev, job, H, False, buf["reduced"], None)
        gathered = exchange_records(buf["reduced"])                       # the one exchange step of the path
        key = ("gather_red", H)
        red2 = pl._bufs.setdefault(key, {"r": torch.empty_like(buf["reduced"])})["r"]
        pl.launch_finalize(gathered, gathered.shape[0], H * pl.rec, pl.rec, jobs_dev, job, H, True, red2, buf["stats"])
        pl.launch_map_derived(map_stats, buf["map_derived"])
        pl.launch_evidence(buf["stats"], pl.srec, map_stats, buf["map_derived"], poses_dev, H, buf["evidence"])
        return LidarScanEvidence(n_bins=pl.n_bins, n_hyp=H, stats=buf["stats"], evidence=buf["evidence"], n_input=n_raw_total,
                                 n_points_cap=pl.n_points_cap)


# ----------------------------------------------------------------------------------------------------------------
# bench helpers (used by bench.py; kept here so the sharded paths they time are the ones above)
# ----------------------------------------------------------------------------------------------------------------
def _time_loop(fn, steps: int, warmup: int, world: int) -> float:
    """Average milliseconds per call: CUDA events on the current stream, barrier + sync both sides, max over ranks."""
    for _ in range(max(3, warmup)):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def bench_points_sharded(g, n_points: int, n_bins: int, tau: float, steps: int, warmup: int, rank: int, world: int,
                         peak_gbs: float, precisions=("fp32c", "fp64")) -> Dict[str, object]:
    """One dense cloud of n_points, sharded by points (strong scaling, as BASELINE.json configs[3] states it).
    Several copies of the input are rotated so that consecutive iterations never re-read an L2-resident cloud."""
    from . import synth
    sc = synth.make_scan(n_points, seed=1234 + 4, dense=True)
    out: Dict[str, object] = {"n_points": n_points, "n_bins": n_bins, "tau": tau, "scaling": "strong",
                              "sharding": "points" if world > 1 else "none"}
    for precision in precisions:
        pl = g.LidarEvidencePipeline(n_bins=n_bins, tau=tau, precision=precision, n_points_cap=n_points, lidar_origin=sc.lidar_origin)
        sh = PointShardedEvidence(pl, rank, world)
        bpp = 20 if precision == "fp32c" else 40
        n_copies = int(np.ceil(140e6 / (n_points * bpp / world))) + 1
        n_copies = max(2, min(n_copies, 64))
        shards = []
        for _ in range(n_copies):
            scan, n_raw = sh.upload_shard(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
            shards.append(scan)
        ms = g.create_empty_map_stats(n_bins)
        poses = torch.zeros(1, 6, dtype=torch.float64, device="cuda")

        def whole(i=0):
            sh.run(shards[i % n_copies], n_points, sc.xi_body[None], poses, ms)

        def k1_only(i=0):
            scan = shards[i % n_copies]
            _, _, (b, e), _ = sh.plan(n_points)
            ctas = pl.plan_ctas(1, max(1, e - b))
            job = pl.make_job(scan, sc.xi_body, n_raw_total=n_points, sel_range=(b, e))
            pl.launch_moments(None, job, 1, ctas, scan.layout, pl._buffers(1, ctas)["partials"])

        ms_whole = _time_loop(whole, steps, warmup, world)
        ms_k1 = _time_loop(k1_only, steps, warmup, world)
        gbs = n_points * bpp / (ms_k1 * 1e-3) / 1e9
        out[precision] = {
            "points_per_sec": n_points / (ms_whole * 1e-3), "scans_per_sec": 1e3 / ms_whole, "ms_per_scan": ms_whole,
            "scan_moments_ms": ms_k1, "pairs_per_sec": n_points * n_bins / (ms_k1 * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak_gbs * world, "unit": "GB/s", "frac": gbs / (peak_gbs * world),
                         "bytes_per_point": bpp},
            "input_copies_rotated": n_copies,
        }
    return out


def bench_hypothesis_sharded(g, n_points: int, cap: int, n_bins: int, n_hyp: int, tau: float, steps: int, warmup: int,
                             rank: int, world: int, precision: str = "fp64") -> Dict[str, object]:
    """configs[2]: n_hyp hypotheses (distinct twists and predicted poses) share one scan; hypotheses are split over ranks."""
    from . import synth
    n_rot = 8
    scans_h = [synth.make_scan(n_points, seed=1234 + 3 + i) for i in range(n_rot)]
    rng = np.random.default_rng(1234 + 3)
    xis = scans_h[0].xi_body[None] * (1 + 0.02 * rng.normal(size=(n_hyp, 6)))
    poses = 0.02 * rng.normal(size=(n_hyp, 6))
    lo, hi = shard_rows(n_hyp, rank, world)
    pl = g.LidarEvidencePipeline(n_bins=n_bins, tau=tau, precision=precision, n_points_cap=cap, lidar_origin=scans_h[0].lidar_origin)
    res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans_h]
    poses_dev = torch.from_numpy(poses[lo:hi].copy()).cuda()
    ms = g.create_empty_map_stats(n_bins)
    jobs_dev = [pl.jobs_to_device([pl.make_job(r, xis[h]) for h in range(lo, hi)]) for r in res]
    gathered = torch.empty((world, hi - lo, g._abi.GCL_EVID_DOUBLES), dtype=torch.float64, device="cuda") if world > 1 else None

    def run(i=0):
        ev = pl.run_resident(res[i % n_rot], xis[lo:hi], poses_dev, ms, jobs_dev=jobs_dev[i % n_rot])
        if world > 1 and (n_hyp % world) == 0:
            dist.all_gather_into_tensor(gathered, ev.evidence)

    ms_step = _time_loop(run, steps, warmup, world)
    return {"n_points": n_points, "n_points_cap": cap, "n_bins": n_bins, "n_hyp": n_hyp, "precision": precision,
            "sharding": "hypotheses" if world > 1 else "none", "ms_per_batch": ms_step,
            "hypothesis_scans_per_sec": n_hyp / (ms_step * 1e-3), "points_per_sec": n_hyp * n_points / (ms_step * 1e-3),
            "note": "the scan (1.2 MB) is shared by all hypotheses and is L2-resident by design; 8 scans are rotated between iterations"}
