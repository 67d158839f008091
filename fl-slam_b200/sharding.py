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

On GPUs the exchange itself is NOT a library collective: ``PeerExchange`` maps every rank's exchange block into every process
(CUDA IPC over NVLink peer memory) and ``gcl_sharded_reduce_finalize`` sums this rank's partial records, stores the result
straight into the peers' blocks, waits for theirs and finalises, in one launch (csrc/gcl_exchange.cu).  The NCCL all-gather
route of round 1 is kept as ``exchange="nccl"`` for comparison.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import os

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


class PeerExchange:
    """This rank's exchange block + the peers' blocks mapped over CUDA IPC (gcl_exchange_create / _connect).  The 64-byte IPC
    handles travel through torch.distributed once, at construction; after that no library collective is involved."""

    def __init__(self, pipeline, rank: int, world: int, slot_bytes: int, group=None):
        from . import _abi
        self._abi, self.pl, self.rank, self.world = _abi, pipeline, rank, world
        self.slot_bytes = int(slot_bytes)
        self.x = C.c_void_p()
        hbuf = (C.c_ubyte * _abi.GCL_EXCHANGE_HANDLE_BYTES)()
        with torch.cuda.device(pipeline.device):
            rc = _abi.lib.gcl_exchange_create(pipeline.handle, world, rank, self.slot_bytes, C.byref(self.x), hbuf)
            if rc != 0:
                raise _abi.GclError(f"gcl_exchange_create failed ({rc})")
            mine = torch.tensor(list(bytes(hbuf)), dtype=torch.uint8, device=pipeline.device)
            allh = torch.empty(world * _abi.GCL_EXCHANGE_HANDLE_BYTES, dtype=torch.uint8, device=pipeline.device)
            if world > 1:
                dist.all_gather_into_tensor(allh, mine, group=group)
            else:
                allh.copy_(mine)
            hb = bytes(allh.cpu().numpy().tobytes())
            self._handles = (C.c_ubyte * len(hb)).from_buffer_copy(hb)
            rc = _abi.lib.gcl_exchange_connect(pipeline.handle, self.x, self._handles)
            if rc != 0:
                raise _abi.GclError(_abi.lib.gcl_exchange_last_error(self.x).decode())
            if world > 1:
                dist.barrier(group=group)          # every rank has mapped every block before the first step

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise self._abi.GclError(self._abi.lib.gcl_exchange_last_error(self.x).decode())

    def error_flag(self) -> int:
        """Non-zero once a step has timed out waiting for a peer (synchronises with the device)."""
        return int(self._abi.lib.gcl_exchange_error_flag(self.x))

    def all_gather(self, src: torch.Tensor, dst: Optional[torch.Tensor]) -> None:
        """src (contiguous, a multiple of 16 bytes) -> dst (world, *src.shape) on every rank; one launch, no library collective."""
        from . import operators as ops
        nbytes = src.numel() * src.element_size()
        self._check(self._abi.lib.gcl_exchange_all_gather(self.pl.handle, self.x, ops._stream(), ops._ptr(src), nbytes,
                                                          None if dst is None else ops._ptr(dst)))

    def close(self) -> None:
        """Collective: every rank unmaps the peers' blocks, and only then frees its own."""
        if self.x:
            torch.cuda.synchronize(self.pl.device)
            if self.world > 1:
                dist.barrier()                      # nobody is still inside a step
            self._abi.lib.gcl_exchange_disconnect(self.pl.handle, self.x)
            if self.world > 1:
                dist.barrier()                      # nobody still has my block mapped
            self._abi.lib.gcl_exchange_destroy(self.pl.handle, self.x)
            self.x = C.c_void_p()


class ScanPipe:
    """Consecutive scans / hypothesis batches in flight (gcl_scan_pipe, csrc/gcl_pipe.cu): one C call per submission; the moment pass
    on one internal stream, the O(B) tail (exchange + finalise + evidence [+ gather]) alternating over ``n_tails`` high-priority
    streams, ``n_sets`` buffer sets.  mode "sum": points sharded (needs a PeerExchange); "gather": hypotheses sharded."""

    def __init__(self, pipeline, px: Optional[PeerExchange], world: int, mode: str, n_sets: int = 4, n_tails: int = 2):
        from . import _abi, operators as ops
        self._abi, self._ops, self.pl, self.px, self.world = _abi, ops, pipeline, px, world
        self.mode = {"sum": 0, "gather": 1}[mode]
        self.n_sets, self.n_tails, self.turn = int(n_sets), int(n_tails), 0
        self.p = C.c_void_p()
        with torch.cuda.device(pipeline.device):
            rc = _abi.lib.gcl_scan_pipe_create(pipeline.handle, None if px is None else px.x, world, self.mode, self.n_sets, self.n_tails,
                                               C.byref(self.p))
        if rc != 0:
            raise _abi.GclError(f"gcl_scan_pipe_create failed ({rc}): {_abi.lib.gcl_last_error(pipeline.handle).decode()}")
        self._keep = [None] * self.n_sets
        self._args: Dict[tuple, tuple] = {}
        self._gathered: Dict[tuple, torch.Tensor] = {}

    def submit(self, jobs_dev: Optional[torch.Tensor], job, n_jobs: int, ctas: int, layout: int, map_stats, poses_dev: torch.Tensor,
               keep=None, wait_for_caller: bool = True):
        """Enqueue one scan (or batch); returns (buffer set dict, gathered tensor or None).  Valid after ``join()``, until n_sets
        further submissions."""
        pl, ops, lib = self.pl, self._ops, self._abi.lib
        i = self.turn % self.n_sets
        self.turn += 1
        key = (n_jobs, ctas, i)
        a = self._args.get(key)
        if a is None:       # the per-set pointers do not change from scan to scan: build the ctypes arguments once
            buf = pl._buffers(n_jobs, ctas, slot=i)
            g = None
            if self.mode == 1 and self.world > 1:
                g = self._gathered.setdefault(key, torch.empty((self.world, n_jobs, self._abi.GCL_EVID_DOUBLES), dtype=torch.float64,
                                                               device=pl.device))
            a = self._args[key] = (buf, g, tuple(ops._ptr(buf[k]) for k in ("partials", "reduced", "stats", "map_derived", "evidence")) +
                                   (ops._ptr(g),))
        buf, g, ptrs = a
        self._keep[i] = (jobs_dev, job, map_stats, poses_dev, keep)      # alive until this set comes round again
        m = map_stats
        rc = lib.gcl_scan_pipe_submit(self.p, ops._stream() if wait_for_caller else None, 1 if wait_for_caller else 0, i,
                                      None if jobs_dev is None else C.c_void_p(jobs_dev.data_ptr()), None if job is None else C.byref(job),
                                      n_jobs, ctas, layout, self._precision(), self._bin_dirs(), pl.n_bins, pl.tau, pl._origin_c, pl.eps_psd,
                                      pl.eps_mass, ops._ptr(m.S_dir), ops._ptr(m.S_dir_scatter), ops._ptr(m.N_dir), ops._ptr(m.N_pos),
                                      ops._ptr(m.sum_p), ops._ptr(m.sum_ppT), ops._ptr(poses_dev), *ptrs)
        if rc != 0:
            raise self._abi.GclError(lib.gcl_scan_pipe_last_error(self.p).decode())
        return buf, g

    def _precision(self) -> int:
        from .pipeline import PRECISIONS
        return PRECISIONS[self.pl.precision]

    def _bin_dirs(self):
        return self._ops._ptr(self.pl.bin_dirs)

    def next_set(self) -> int:
        """The buffer set the next submission writes."""
        return self.turn % self.n_sets

    def wait_set(self, i: int) -> None:
        """Order the caller's CURRENT stream after the last submission that used set ``i`` (gcl_scan_pipe_wait_set): what the caller
        enqueues next on that stream may overwrite the staging buffers and host outputs tied to the set."""
        rc = self._abi.lib.gcl_scan_pipe_wait_set(self.p, int(i), self._ops._stream())
        if rc != 0:
            raise self._abi.GclError(self._abi.lib.gcl_scan_pipe_last_error(self.p).decode())

    def set_readback(self, i: int, stats_host: Optional[torch.Tensor], evidence_host: Optional[torch.Tensor]) -> None:
        """Page-locked host tensors every later submission on set ``i`` fills at the end of its tail (gcl_scan_pipe_set_readback)."""
        for t in (stats_host, evidence_host):
            if t is not None and not (t.is_pinned() and t.is_contiguous() and t.dtype == torch.float64):
                raise ValueError("read-back buffers must be contiguous page-locked float64 tensors")
        self._readback = getattr(self, "_readback", {})
        self._readback[int(i)] = (stats_host, evidence_host)        # kept alive with the pipe
        rc = self._abi.lib.gcl_scan_pipe_set_readback(self.p, int(i), None if stats_host is None else C.c_void_p(stats_host.data_ptr()),
                                                      None if evidence_host is None else C.c_void_p(evidence_host.data_ptr()))
        if rc != 0:
            raise self._abi.GclError("gcl_scan_pipe_set_readback: bad arguments")

    def join(self) -> None:
        """Order the caller's current stream after everything submitted."""
        rc = self._abi.lib.gcl_scan_pipe_join(self.p, self._ops._stream())
        if rc != 0:
            raise self._abi.GclError(self._abi.lib.gcl_scan_pipe_last_error(self.p).decode())

    def close(self) -> None:
        if self.p:
            self._abi.lib.gcl_scan_pipe_destroy(self.p)       # synchronises its streams
            self.p = C.c_void_p()
            self._keep = [None] * self.n_sets


class PointShardedEvidence:
    """configs[3]/[4]: one scan, points sharded over the ranks of the default process group.

    exchange = "peer": the fused NVLink exchange + finalisation kernel (one launch between scan_moments and the chain step);
    "nccl": reduce -> all_gather_into_tensor -> finalize (round 1); "auto": peer on CUDA with world > 1 ranks, else none."""

    def __init__(self, pipeline, rank: int, world: int, exchange: str = "auto", max_jobs: int = 1):
        self.pl, self.rank, self.world = pipeline, rank, world
        if exchange == "auto":
            exchange = "peer"
        self.exchange = exchange
        self.px: Optional[PeerExchange] = None
        if exchange == "peer":
            self.px = PeerExchange(pipeline, rank, world, max_jobs * pipeline.rec * 8)

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
        if self.px is not None:
            # the one exchange step of the path, fused: local fixed-order sum -> peers' blocks over NVLink -> flags -> finalise
            from . import _abi, operators as ops
            if H * pl.rec * 8 > self.px.slot_bytes:
                raise ValueError("more hypotheses than the exchange was sized for (max_jobs)")
            self.px._check(_abi.lib.gcl_sharded_reduce_finalize(
                pl.handle, self.px.x, ops._stream(), ops._ptr(buf["partials"]), ctas, pl.rec, pl.rec * ctas,
                None if jobs_dev is None else C.c_void_p(jobs_dev.data_ptr()), None if job is None else C.byref(job), H, pl.n_bins,
                ops._ptr(pl.bin_dirs), pl.tau, pl._origin_c, pl.eps_psd, pl.eps_mass, ops._ptr(buf["reduced"]), ops._ptr(buf["stats"])))
        else:
            pl.launch_finalize(buf["partials"], ctas, pl.rec, pl.rec * ctas, jobs_dev, job, H, False, buf["reduced"], None)
            gathered = exchange_records(buf["reduced"])                   # NCCL all-gather (round 1 route)
            key = ("gather_red", H)
            red2 = pl._bufs.setdefault(key, {"r": torch.empty_like(buf["reduced"])})["r"]
            pl.launch_finalize(gathered, gathered.shape[0], H * pl.rec, pl.rec, jobs_dev, job, H, True, red2, buf["stats"])
        # map-derived statistics + pose evidence: the single-launch chain step (read-only map, no IW state)
        pl.launch_chain(buf["stats"], 1, H, map_stats, buf["map_derived"], poses_dev, None, buf["evidence"])
        return LidarScanEvidence(n_bins=pl.n_bins, n_hyp=H, stats=buf["stats"], evidence=buf["evidence"], n_input=n_raw_total,
                                 n_points_cap=pl.n_points_cap)

    def _ensure_overlapped(self, reserve_sms: int, n_tails: int):
        """The scan pipe for this (reserve_sms, n_tails), created on first use."""
        from . import _abi
        pl = self.pl
        if self.px is None:
            raise ValueError("run_overlapped needs the peer exchange (exchange='peer')")
        ov = getattr(self, "_ov", None)
        if ov is None or ov["key"] != (reserve_sms, n_tails):
            if ov is not None:
                self.end_overlapped()
            sms = int(_abi.lib.gcl_device_sm_count(pl.handle))
            ov = self._ov = {"key": (reserve_sms, n_tails), "sms": sms, "pipe": ScanPipe(pl, self.px, self.world, "sum", n_sets=4, n_tails=n_tails)}
            # grid cap of the exchange kernel: its CTAs must find room beside the next scan's moment pass -- as many as SMs are left free
            # (the two tails rarely sum their partial records at the same moment; measured at 2 GPUs, fp32c: cap 1 / 2 / 4 / 8 with 2 SMs
            # free = 0.138 / 0.135 / 0.163 / 0.139 ms per scan); GCL_EXCHANGE_CTAS overrides
            cap = int(os.environ.get("GCL_EXCHANGE_CTAS", "0")) or max(1, reserve_sms)
            _abi.lib.gcl_exchange_set_max_ctas(self.px.x, cap)
        return ov

    # ---- consecutive scans pipelined: the exchange + evidence step of scan k under the moment pass of scan k + 1 -------------------
    def run_overlapped(self, scan, n_raw_total: int, xi_bodies, poses_dev, map_stats, reserve_sms: int = 8, n_tails: int = 2, job=None):
        """Same result as ``run``; the moment pass goes on one internal stream with ``reserve_sms`` SMs left free, the fused exchange +
        finalisation (grid capped to the free SMs) and the evidence step on high-priority streams of their own -- so the O(B) tail of
        scan k (60-70 us, latency-bound) runs beside the moment pass of scan k + 1 instead of after it, and with ``n_tails`` = 2 beside
        the tail of scan k + 1 as well (needed once the shard's moment pass is shorter than a tail: 8 GPUs, fp32c).  One C call per scan
        (gcl_scan_pipe_submit).  Results rotate through four buffer sets (valid until four calls later); ``join_overlapped()`` orders the
        caller's stream after everything submitted.  ``job``: a gcl_job built earlier with ``make_sharded_job`` (saves its cost)."""
        from .pipeline import LidarScanEvidence
        pl = self.pl
        ov = self._ensure_overlapped(reserve_sms, n_tails)
        jobs_dev = None
        if job is None:
            xi = np.asarray(xi_bodies, dtype=np.float64).reshape(-1, 6)
            H = xi.shape[0]
            _, _, (b, e), _ = self.plan(n_raw_total)
            jobs = [pl.make_job(scan, xi[h], n_raw_total=n_raw_total, sel_range=(b, e)) for h in range(H)]
            job, jobs_dev = (jobs[0], None) if H == 1 else (None, pl.jobs_to_device(jobs))
        else:
            H = 1
        ck = (n_raw_total, H)
        ctas = ov.get(ck)
        if ctas is None:
            _, _, (b, e), _ = self.plan(n_raw_total)
            ctas = ov[ck] = max(1, min(pl.plan_ctas(H, max(1, e - b)), (ov["sms"] - reserve_sms) // max(1, H)))
        buf, _ = ov["pipe"].submit(jobs_dev, job, H, ctas, scan.layout, map_stats, poses_dev, keep=scan)
        return LidarScanEvidence(n_bins=pl.n_bins, n_hyp=H, stats=buf["stats"], evidence=buf["evidence"], n_input=n_raw_total,
                                 n_points_cap=pl.n_points_cap)

    def make_sharded_job(self, scan, n_raw_total: int, xi_body):
        """The gcl_job of this rank's share of ``scan`` (for ``run_overlapped(job=...)``)."""
        _, _, (b, e), _ = self.plan(n_raw_total)
        return self.pl.make_job(scan, np.asarray(xi_body, dtype=np.float64).reshape(6), n_raw_total=n_raw_total, sel_range=(b, e))

    def join_overlapped(self) -> None:
        ov = getattr(self, "_ov", None)
        if ov is not None:
            ov["pipe"].join()

    def end_overlapped(self) -> None:
        """Back to one scan at a time (the exchange kernel's grid is no longer capped)."""
        from . import _abi
        ov = getattr(self, "_ov", None)
        if ov is not None:
            ov["pipe"].join()
            ov["pipe"].close()
        if self.px is not None:
            _abi.lib.gcl_exchange_set_max_ctas(self.px.x, 0)
        self._ov = None

    # ---- the same scan from HOST memory: this rank's shard crosses PCIe inside the call --------------------------------------
    def host_stage(self, points, timestamps, weights, scan_start_time: float, scan_end_time: float):
        """Page-locked copies of this rank's raw rows + two device staging shards (so that the upload of scan k + 1 can run on a
        copy stream under the kernels of scan k)."""
        n_raw = int(np.asarray(points).shape[0])
        _, _, _, (lo, hi) = self.plan(n_raw)
        pin = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).pin_memory() for a in (points, timestamps, weights)]
        dev = [self.pl.upload_scan(points[lo:hi], timestamps[lo:hi], weights[lo:hi], scan_start_time, scan_end_time, packed=False, row_offset=lo)
               for _ in range(2)]
        st = {"pin": pin, "dev": dev, "n_raw": n_raw, "copy": torch.cuda.Stream(device=self.pl.device), "ready": [None, None], "free": [None, None],
              "out": [torch.empty((1, self.pl.srec), dtype=torch.float64).pin_memory(), torch.empty((1, 640), dtype=torch.float64).pin_memory()],
              "turn": 0, "h2d_bytes": int(sum(t.numel() * t.element_size() for t in pin)), "d2h_bytes": int((self.pl.srec + 640) * 8)}
        return st

    def enqueue_host_upload(self, st) -> None:
        """H2D of the shard into the staging buffer of the next turn, on the copy stream."""
        i = st["turn"] & 1
        cs = st["copy"]
        if st["free"][i] is not None:
            cs.wait_event(st["free"][i])                     # the kernels that read this staging buffer two turns ago are done
        with torch.cuda.stream(cs):
            d = st["dev"][i]
            d.points.copy_(st["pin"][0], non_blocking=True); d.timestamps.copy_(st["pin"][1], non_blocking=True)
            d.weights.copy_(st["pin"][2], non_blocking=True)
            st["ready"][i] = cs.record_event()

    def run_host(self, st, xi_bodies, poses_dev, map_stats):
        """One scan end to end: (upload enqueued by enqueue_host_upload) -> sharded evidence -> D2H of this rank's stats and
        evidence records into page-locked memory.  The caller synchronises."""
        i = st["turn"] & 1
        st["turn"] += 1
        cur = torch.cuda.current_stream(self.pl.device)
        cur.wait_event(st["ready"][i])
        if self.pl.precision == "fp32c":                      # the packed 20 B/point layout is produced on the device from the upload
            d = st["dev"][i]
            scan = self.pl.upload_scan(d.points, d.timestamps, d.weights, d.scan_start_time, d.scan_end_time, packed=True, row_offset=d.row_offset)
        else:
            scan = st["dev"][i]
        ev = self.run(scan, st["n_raw"], xi_bodies, poses_dev, map_stats)
        st["out"][0].copy_(ev.stats, non_blocking=True); st["out"][1].copy_(ev.evidence, non_blocking=True)
        st["free"][i] = cur.record_event()
        return ev


    # ---- scans from HOST memory, consecutive scans in flight: upload of scan k + 1 under the kernels of scan k, one C call per scan ----
    def host_stage_pipelined(self, points, timestamps, weights, scan_start_time: float, scan_end_time: float, n_hyp: int = 1,
                             reserve_sms: int = 8, n_tails: int = 2):
        """Page-locked copies of this rank's raw rows; per buffer set of the scan pipe (four) one device staging shard and one pair of
        page-locked result buffers, named to the pipe as the read-back targets of that set."""
        n_raw = int(np.asarray(points).shape[0])
        _, _, _, (lo, hi) = self.plan(n_raw)
        pin = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).pin_memory() for a in (points, timestamps, weights)]
        pipe = self._ensure_overlapped(reserve_sms, n_tails)["pipe"]
        dev = [self.pl.upload_scan(points[lo:hi], timestamps[lo:hi], weights[lo:hi], scan_start_time, scan_end_time, packed=False, row_offset=lo)
               for _ in range(pipe.n_sets)]
        out = [(torch.empty((n_hyp, self.pl.srec), dtype=torch.float64).pin_memory(), torch.empty((n_hyp, 640), dtype=torch.float64).pin_memory())
               for _ in range(pipe.n_sets)]
        for k in range(pipe.n_sets):
            pipe.set_readback(k, out[k][0], out[k][1])
        cs = torch.cuda.Stream(device=self.pl.device)
        cs.wait_stream(torch.cuda.current_stream(self.pl.device))    # the staging shards, poses and map the caller has enqueued so far
        return {"pin": pin, "dev": dev, "out": out, "n_raw": n_raw, "copy": cs, "key": (reserve_sms, n_tails),
                "h2d_bytes": int(sum(t.numel() * t.element_size() for t in pin)), "d2h_bytes": int(n_hyp * (self.pl.srec + 640) * 8)}

    def run_host_pipelined(self, st, xi_bodies, poses_dev, map_stats):
        """One scan from page-locked host arrays through the scan pipe: H2D of this rank's rows on the copy stream (once the scan that
        used the same set four calls ago has finished), then ONE submission whose moment pass waits for the upload and whose tail ends
        with the D2H of the stats and evidence records into ``st["out"][set]``.  Nothing here blocks the host: the upload of scan k + 1
        runs under the kernels of scan k.  ``join_overlapped()`` + a synchronisation make the results visible; same results as
        ``run_host`` (the same launches).  Returns (evidence view on the device, set index)."""
        pl = self.pl
        reserve_sms, n_tails = st["key"]
        pipe = self._ensure_overlapped(reserve_sms, n_tails)["pipe"]
        with torch.cuda.stream(st["copy"]):
            i = pipe.next_set()
            pipe.wait_set(i)
            d = st["dev"][i]
            d.points.copy_(st["pin"][0], non_blocking=True); d.timestamps.copy_(st["pin"][1], non_blocking=True)
            d.weights.copy_(st["pin"][2], non_blocking=True)
            if pl.precision == "fp32c":                      # the packed 20 B/point layout is produced on the device from the upload
                scan = pl.upload_scan(d.points, d.timestamps, d.weights, d.scan_start_time, d.scan_end_time, packed=True, row_offset=d.row_offset)
            else:
                scan = d
            ev = self.run_overlapped(scan, st["n_raw"], xi_bodies, poses_dev, map_stats, reserve_sms=reserve_sms, n_tails=n_tails)
        return ev, i

    def end_host_pipelined(self, st) -> None:
        """Forget the read-back targets (the pipe outlives ``st``)."""
        ov = getattr(self, "_ov", None)
        if ov is not None:
            ov["pipe"].join()
            torch.cuda.synchronize(self.pl.device)
            for k in range(ov["pipe"].n_sets):
                ov["pipe"].set_readback(k, None, None)


# ----------------------------------------------------------------------------------------------------------------
# bench helpers (used by bench.py; kept here so the sharded paths they time are the ones above)
# ----------------------------------------------------------------------------------------------------------------
def _max_over_ranks(x: float, world: int) -> float:
    if world > 1:
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        x = float(t.item())
    return x


def _time_steps(fn, steps: int, warmup: int, world: int):
    """fn(step_index) enqueues ONE step.  W untimed steps, then EXACTLY `steps` steps between a barrier + synchronize on both sides,
    CUDA events on the current stream around every step.  Returns (ms per step over the whole region -- max over ranks --, per-step
    times of this rank)."""
    for k in range(max(3, warmup)):
        fn(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for k in range(steps):
        fn(k)
        evs[k + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total = evs[0].elapsed_time(evs[steps])
    per = [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]
    return _max_over_ranks(total, world) / steps, per


def _stats_arrays(pl, stats: torch.Tensor) -> Dict[str, torch.Tensor]:
    B = pl.n_bins
    s = stats.reshape(-1)
    return {"N": s[:B], "s_dir": s[B:4 * B], "S_dir_scatter": s[4 * B:13 * B], "p_bar": s[13 * B:16 * B], "Sigma_p": s[16 * B:25 * B],
            "kappa": s[25 * B:26 * B], "sum_p": s[26 * B:29 * B], "sum_ppT": s[29 * B:38 * B]}


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    d = float((a - b).abs().max())
    m = float(b.abs().max())
    return d / (m if m > 0 else 1.0)


def _bits_identical_across_ranks(t: torch.Tensor, world: int) -> bool:
    """Every rank holds the very same bits: compare a 64-bit wrap-around checksum of the int64 view and the raw first / last words."""
    v = t.contiguous().view(torch.int64).reshape(-1)
    sig = torch.stack([v.sum(), v[0], v[-1], (v * torch.arange(1, v.numel() + 1, device=v.device, dtype=torch.int64)).sum()])
    if world == 1:
        return True
    allsig = torch.empty((world, 4), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(allsig, sig)
    return bool((allsig == allsig[0:1]).all())


def bench_points_sharded(g, n_points: int, n_bins: int, tau: float, steps: int, warmup: int, rank: int, world: int,
                         peak_gbs: float, precision: str = "fp64", scans_per_step: int = 64, exchange: str = "peer",
                         with_e2e: bool = True, reserve_sms: Optional[int] = None, n_tails: Optional[int] = None) -> Dict[str, object]:
    """BASELINE.json configs[3]: one dense cloud of n_points, sharded by points over the ranks (strong scaling).  One "step" =
    `scans_per_step` scans (each: scan_moments on this rank's rows, the fused NVLink exchange + finalisation, the evidence step);
    several copies of the input are rotated so that consecutive scans never re-read an L2-resident cloud."""
    from . import synth
    if reserve_sms is None:
        # SMs left to the exchange + evidence tail of the previous scan: enough that the tail (latency-bound, ~50-70 us) finishes
        # under the moment pass of the next scan, which shrinks with the number of ranks
        # (one GPU: the moment pass of the whole cloud is ~10 x longer than a tail, one SM is enough for it -- tools/tail_probe.py 1048576 256:
        # 848 against 855 us per fp64 scan with 1 / 2 SMs left free, 248 against 249 in fp32c)
        reserve_sms = {1: 1, 2: 2, 4: 4}.get(world, 6)
    if n_tails is None:
        # two tails in flight: a tail (exchange capped to the few SMs left free + finalise + evidence chain) takes 60-160 us, which is
        # longer than the fp32c moment pass of a shard from 2 GPUs on
        n_tails = int(os.environ.get("GCL_BENCH_TAILS", "0")) or 2
    sc = synth.make_scan(n_points, seed=1234 + 4, dense=True)
    pl = g.LidarEvidencePipeline(n_bins=n_bins, tau=tau, precision=precision, n_points_cap=n_points, lidar_origin=sc.lidar_origin)
    sh = PointShardedEvidence(pl, rank, world, exchange=exchange)
    bpp = 20 if precision == "fp32c" else 40
    n_copies = int(np.ceil(300e6 / (n_points * bpp / world))) + 1       # > 2 x the 126 MB L2 per rank
    n_copies = max(2, min(n_copies, 96))
    shards = [sh.upload_shard(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)[0] for _ in range(n_copies)]
    rs = synth.known_map_transform(4)
    ms = g.create_empty_map_stats(n_bins)
    poses = torch.zeros(1, 6, dtype=torch.float64, device="cuda")
    xi = sc.xi_body[None]
    _, n_sel, (b, e), _ = sh.plan(n_points)
    ctas = pl.plan_ctas(1, max(1, e - b))

    # a map to register against: this cloud's own bin statistics moved by a known transform, from one sharded pass
    ev0 = sh.run(shards[0], n_points, xi, poses, ms)
    torch.cuda.synchronize()
    sa = {k: v.detach().cpu().numpy() for k, v in _stats_arrays(pl, ev0.stats[0]).items()}
    raw = {"N": sa["N"], "s_dir": sa["s_dir"].reshape(-1, 3), "S_dir_scatter": sa["S_dir_scatter"].reshape(-1, 3, 3),
           "sum_p": sa["sum_p"].reshape(-1, 3), "sum_ppT": sa["sum_ppT"].reshape(-1, 3, 3)}
    ms = g.MapBinStats.from_arrays(synth.map_stats_from_scan_sums(raw, rs[0], rs[1], seed=4))

    it = [0]
    overlapped = sh.px is not None and reserve_sms > 0
    ctas_run = max(1, min(ctas, int(g._abi.lib.gcl_device_sm_count(pl.handle)) - reserve_sms)) if overlapped else ctas

    jobs_pre = [sh.make_sharded_job(s_, n_points, sc.xi_body) for s_ in shards]     # one gcl_job per resident copy, built once

    def step(_k):
        # consecutive scans are pipelined: the exchange + evidence tail of scan k (three small launches' worth of latency on a few SMs)
        # runs under the moment pass of scan k + 1; one C call per scan (gcl_scan_pipe_submit)
        for _ in range(scans_per_step):
            if overlapped:
                c = it[0] % n_copies
                sh.run_overlapped(shards[c], n_points, xi, poses, ms, reserve_sms=reserve_sms, n_tails=n_tails, job=jobs_pre[c])
            else:
                sh.run(shards[it[0] % n_copies], n_points, xi, poses, ms, ctas=ctas)
            it[0] += 1
        if overlapped:
            sh.join_overlapped()

    def step_serial(_k):
        for _ in range(scans_per_step):
            sh.run(shards[it[0] % n_copies], n_points, xi, poses, ms, ctas=ctas)
            it[0] += 1

    def k1_step(_k):
        for _ in range(scans_per_step):
            scan = shards[it[0] % n_copies]
            it[0] += 1
            job = pl.make_job(scan, sc.xi_body, n_raw_total=n_points, sel_range=(b, e))
            pl.launch_moments(None, job, 1, ctas_run, scan.layout, pl._buffers(1, ctas_run)["partials"])

    l0 = g._abi.launch_count(pl.device_index)
    step(0)
    torch.cuda.synchronize()
    launches_per_step = g._abi.launch_count(pl.device_index) - l0
    ms_step, per = _time_steps(step, steps, warmup, world)
    if overlapped:
        sh.end_overlapped()
    ms_serial_step, _ = _time_steps(step_serial, max(2, steps // 2), 1, world)
    ms_k1_step, _ = _time_steps(k1_step, max(2, steps // 2), 1, world)
    ms_scan, ms_k1 = ms_step / scans_per_step, ms_k1_step / scans_per_step
    shard_bytes = (e - b) * bpp                                   # algorithmic bytes of one scan_moments launch on this rank
    gbs = shard_bytes / (ms_k1 * 1e-3) / 1e9
    per = np.sort(np.asarray(per)) / scans_per_step

    # ---- parity of the sharded CUDA path: against the SAME job unsharded on this GPU, and bit-identity across ranks ------------
    ev = sh.run(shards[0], n_points, xi, poses, ms, ctas=ctas)
    st_sh, ev_sh = ev.stats.clone(), ev.evidence.clone()
    full = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
    evf = pl.run_resident(full, xi, poses, ms)
    a_sh, a_f = _stats_arrays(pl, st_sh[0]), _stats_arrays(pl, evf.stats[0])
    parity = {k: _rel(a_sh[k], a_f[k]) for k in a_f}
    R_sh, R_f = ev_sh[0, 0:9], evf.evidence[0, 0:9]
    parity_rot = float((R_sh - R_f).norm() / np.sqrt(2.0))
    parity_t = _rel(ev_sh[0, 71:74], evf.evidence[0, 71:74])
    identical = _bits_identical_across_ranks(torch.cat([st_sh.reshape(-1), ev_sh.reshape(-1)]), world)
    err_flag = sh.px.error_flag() if sh.px is not None else 0
    del full

    tc_mode = int(os.environ.get("GCL_K1_TC", "1"))
    on_tc = precision == "fp32c" and n_bins <= 256 and tc_mode != 0 and (tc_mode == 2 or n_bins >= 128)     # tc_path_supported (gcl_scan_tc.cu)
    k1_name = "scan_moments_tc_kernel (tcgen05)" if on_tc else "scan_moments_fused_kernel"
    out: Dict[str, object] = {
        "n_points": n_points, "n_bins": n_bins, "tau": tau, "precision": precision, "scaling": "strong",
        "sharding": "points" if world > 1 else "none", "exchange": sh.exchange if world > 1 else "none (1 rank: same fused kernel, no peers)",
        "scans_per_step": scans_per_step, "ms_per_step": ms_step, "ms_per_scan": ms_scan, "ms_per_scan_median": float(per[len(per) // 2]),
        "ms_per_scan_best": float(per[0]), "points_per_sec": n_points / (ms_scan * 1e-3), "scans_per_sec": 1e3 / ms_scan,
        "scan_moments_ms": ms_k1, "rest_of_scan_ms": max(0.0, ms_scan - ms_k1), "pairs_per_sec": n_points * n_bins / (ms_k1 * 1e-3),
        "ms_per_scan_one_at_a_time": ms_serial_step / scans_per_step,
        "pipelining": (f"consecutive scans in flight (gcl_scan_pipe, one call per scan): scan_moments on {ctas_run} CTAs on one stream, the exchange + "
                       f"evidence tails of the previous scans on {n_tails} high-priority stream(s) and the {reserve_sms} SMs left free") if overlapped
                      else "none (one scan at a time)",
        "launches_per_step": int(launches_per_step), "launches_per_scan": launches_per_step / scans_per_step, "ctas_per_rank": ctas_run,
        "rows_per_rank": int(e - b), "input_copies_rotated": n_copies,
        "roofline": {"bound": "hbm", "kernel": k1_name, "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
                     "ms_per_launch": ms_k1, "algorithmic_bytes_per_launch": float(shard_bytes), "bytes_per_point": bpp,
                     "share_of_step": ms_k1 / ms_scan, "launches_per_step": scans_per_step, "per": "GPU (each rank's launch covers its own rows)"},
        "parity": {"vs_unsharded_max_rel": parity, "vs_unsharded_R_mf_rad": parity_rot, "vs_unsharded_t_wls_rel": parity_t,
                   "ranks_bit_identical": bool(identical), "exchange_timeouts": int(err_flag),
                   "note": "the sharded job against the same job run unsharded on this GPU (sums linear in the points: only the grouping of "
                           "the fp64 additions differs); stats + evidence records compared bit for bit across ranks"},
    }
    if with_e2e:
        st = sh.host_stage(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
        sh.enqueue_host_upload(st)

        def e2e_step(_k):
            for _ in range(scans_per_step):
                # the next scan's upload (copy stream) runs under this scan's kernels
                cur = st["turn"]
                st["turn"] = cur + 1
                sh.enqueue_host_upload(st)
                st["turn"] = cur
                sh.run_host(st, xi, poses, ms)
            torch.cuda.synchronize()                         # the step's results are in host memory when the step ends

        t_e2e, _ = _time_steps(e2e_step, max(2, steps // 2), 1, world)
        note = ("every scan: H2D of this rank's raw rows (points, timestamps, weights, float64) from page-locked host arrays on a copy "
                "stream, the sharded evidence pass, D2H of the stats + evidence records; the upload of scan k + 1 overlaps the "
                "kernels of scan k; host synchronisation once per step; bytes are summed over the ranks")
        t_one, e2e_check = t_e2e, None
        if overlapped:
            # the same through the scan pipe (run_host_pipelined): one C call per scan, uploads ordered by the pipe's own "set free" events,
            # the D2H of the records at the end of every tail -- consecutive scans in flight exactly as in the resident measurement
            del st
            stp = sh.host_stage_pipelined(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time, reserve_sms=reserve_sms,
                                          n_tails=n_tails)

            def e2e_step_pipe(_k):
                for _ in range(scans_per_step):
                    sh.run_host_pipelined(stp, xi, poses, ms)
                sh.join_overlapped()
                torch.cuda.synchronize()

            t_e2e, _ = _time_steps(e2e_step_pipe, max(2, steps // 2), 1, world)
            # every set's page-locked records equal, bit for bit, what the same submission gives on the resident copy of the cloud
            ev_r = sh.run_overlapped(shards[0], n_points, xi, poses, ms, reserve_sms=reserve_sms, n_tails=n_tails, job=jobs_pre[0])
            sh.join_overlapped()
            torch.cuda.synchronize()
            ref_stats, ref_evid = ev_r.stats.cpu(), ev_r.evidence.cpu()
            e2e_check = {"readback_equals_resident_run_bit_for_bit": bool(all(torch.equal(o[0], ref_stats) and torch.equal(o[1], ref_evid)
                                                                              for o in stp["out"]))}
            sh.end_host_pipelined(stp)
            st = stp
            note = ("every scan: H2D of this rank's raw rows (points, timestamps, weights, float64) from page-locked host arrays on a copy "
                    "stream, ONE gcl_scan_pipe_submit whose moment pass waits for the upload and whose tail ends with the D2H of the stats + "
                    "evidence records into page-locked memory; consecutive scans in flight as in the resident measurement (uploads under the "
                    "kernels of the scans before, four staging sets); host synchronisation once per step; bytes are summed over the ranks")
        out["e2e"] = {"value": n_points / (t_e2e / scans_per_step * 1e-3), "unit": "points/s",
                      "h2d_bytes_per_step": int(st["h2d_bytes"] * scans_per_step * world), "d2h_bytes_per_step": int(st["d2h_bytes"] * scans_per_step * world),
                      "ms_per_scan": t_e2e / scans_per_step, "ms_per_scan_one_scan_at_a_time": t_one / scans_per_step,
                      "h2d_gbs_per_gpu": st["h2d_bytes"] / (t_e2e / scans_per_step * 1e-3) * 1e-9, "note": note}
        if e2e_check:
            out["e2e"].update(e2e_check)
    if sh.px is not None:
        sh.px.close()
    return out


def bench_hypothesis_sharded(g, n_points: int, cap: int, n_bins: int, n_hyp: int, tau: float, steps: int, warmup: int,
                             rank: int, world: int, precision: str = "fp64", batches_per_step: int = 16) -> Dict[str, object]:
    """configs[2]: n_hyp hypotheses (distinct twists and predicted poses) share one scan; hypotheses are split over ranks, no
    data-path collective; the evidence records are gathered at the end through the peer exchange (one launch)."""
    from . import synth
    n_rot = 8
    scans_h = [synth.make_scan(n_points, seed=1234 + 3 + i) for i in range(n_rot)]
    rng = np.random.default_rng(1234 + 3)
    xis = scans_h[0].xi_body[None] * (1 + 0.02 * rng.normal(size=(n_hyp, 6)))
    poses = 0.02 * rng.normal(size=(n_hyp, 6))
    lo, hi = shard_rows(n_hyp, rank, world)
    even = (n_hyp % world) == 0
    pl = g.LidarEvidencePipeline(n_bins=n_bins, tau=tau, precision=precision, n_points_cap=cap, lidar_origin=scans_h[0].lidar_origin)
    res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans_h]
    poses_dev = torch.from_numpy(poses[lo:hi].copy()).cuda()
    ms = g.create_empty_map_stats(n_bins)
    ev0 = pl.run_resident(res[0], xis[0:1], torch.zeros(1, 6, dtype=torch.float64, device="cuda"), ms)
    torch.cuda.synchronize()
    sa = {k: v.detach().cpu().numpy() for k, v in _stats_arrays(pl, ev0.stats[0]).items()}
    raw = {"N": sa["N"], "s_dir": sa["s_dir"].reshape(-1, 3), "S_dir_scatter": sa["S_dir_scatter"].reshape(-1, 3, 3),
           "sum_p": sa["sum_p"].reshape(-1, 3), "sum_ppT": sa["sum_ppT"].reshape(-1, 3, 3)}
    ms = g.MapBinStats.from_arrays(synth.map_stats_from_scan_sums(raw, *synth.known_map_transform(3), seed=3))
    jobs_dev = [pl.jobs_to_device([pl.make_job(r, xis[h]) for h in range(lo, hi)]) for r in res]
    E = g._abi.GCL_EVID_DOUBLES
    px = PeerExchange(pl, rank, world, (hi - lo) * E * 8) if (world > 1 and even) else None
    gathered = torch.empty((world, hi - lo, E), dtype=torch.float64, device="cuda") if px is not None else None
    it = [0]

    def one(i):
        ev = pl.run_resident(res[i % n_rot], xis[lo:hi], poses_dev, ms, jobs_dev=jobs_dev[i % n_rot])
        if px is not None:
            px.all_gather(ev.evidence, gathered)
        return ev

    def step_serial(_k):
        for _ in range(batches_per_step):
            one(it[0])
            it[0] += 1

    # consecutive batches pipelined, as the point-sharded scans are: the moment pass of batch k + 1 (one internal stream, a few SMs left
    # free) runs beside the O(B) tail of batch k -- reduce_finalize (one CTA per hypothesis), map-derived statistics, pose evidence, the
    # gather of the records -- on a second, high-priority stream.  Two buffer sets alternate.
    from . import _abi
    Hl = hi - lo
    sms = int(_abi.lib.gcl_device_sm_count(pl.handle))
    reserve = 8
    _, n_sel = pl.selected_rows(res[0].n_raw, cap)
    ctas_p = max(1, min(pl.plan_ctas(Hl, n_sel), (sms - reserve) // max(1, Hl)))
    n_tails = int(os.environ.get("GCL_BENCH_TAILS", "0")) or 2
    pipe = ScanPipe(pl, px, world, "gather", n_sets=4, n_tails=n_tails)
    ov = {"last": None}

    def one_pipelined(i):
        buf, g_ = pipe.submit(jobs_dev[i % n_rot], None, Hl, ctas_p, res[i % n_rot].layout, ms, poses_dev, wait_for_caller=False)
        ov["last"] = g_
        return buf

    def step(_k):
        for _ in range(batches_per_step):
            one_pipelined(it[0])
            it[0] += 1
        pipe.join()

    ms_step, per = _time_steps(step, steps, warmup, world)
    ms_serial_step, _ = _time_steps(step_serial, max(2, steps // 2), 1, world)
    ms_batch = ms_step / batches_per_step
    # the pipelined batch equals the batch run on one stream, bit for bit (same kernels, same grids per job are NOT guaranteed: the CTA
    # count per job differs when SMs are reserved, which regroups the fp64 additions -- so compare to 1e-12, not bitwise)
    torch.cuda.synchronize()
    bufp = one_pipelined(0)
    pipe.join()
    torch.cuda.synchronize()
    ev_p = bufp["evidence"].clone()
    gathered_p = ov["last"].clone() if ov["last"] is not None else None
    pipe.close()
    # parity: hypothesis h of the batch against the same hypothesis run alone; gathered records bit-identical on every rank
    ev = one(0)
    torch.cuda.synchronize()
    worst = 0.0
    for h in (lo, hi - 1):
        single = pl.run_resident(res[0], xis[h:h + 1], poses_dev[h - lo:h - lo + 1], ms)
        worst = max(worst, float((single.evidence[0] - ev.evidence[h - lo]).abs().max() / max(1e-300, float(single.evidence[0].abs().max()))))
    identical = _bits_identical_across_ranks(gathered.reshape(-1), world) if gathered is not None else True
    if gathered_p is not None:      # the records gathered by the pipelined route: same on every rank, and this rank's own block is its evidence
        identical = identical and _bits_identical_across_ranks(gathered_p.reshape(-1), world) and bool((gathered_p[rank] == ev_p).all())
    pipe_vs_serial = float((ev_p - ev.evidence).abs().max() / max(1e-300, float(ev.evidence.abs().max())))
    out = {"n_points": n_points, "n_points_cap": cap, "n_bins": n_bins, "n_hyp": n_hyp, "precision": precision,
           "sharding": "hypotheses" if world > 1 else "none", "ms_per_batch": ms_batch, "batches_per_step": batches_per_step,
           "ms_per_batch_one_at_a_time": ms_serial_step / batches_per_step,
           "pipelining": f"consecutive batches in flight (gcl_scan_pipe, one call per batch): scan_moments on {ctas_p} CTAs per hypothesis x {Hl} "
                         f"hypotheses on one stream, the finalise + evidence + gather tails of the previous batches on {n_tails} high-priority "
                         f"stream(s) and the {reserve} SMs left free; pipelined vs one-at-a-time evidence max rel {pipe_vs_serial:.1e}",
           "hypothesis_scans_per_sec": n_hyp / (ms_batch * 1e-3), "points_per_sec": n_hyp * n_points / (ms_batch * 1e-3),
           "result_gather": "peer exchange (one launch, NVLink stores + flags)" if px is not None else "none",
           "parity": {"batch_vs_single_hypothesis_max_rel": worst, "gathered_bit_identical_across_ranks": bool(identical),
                      "exchange_timeouts": int(px.error_flag()) if px is not None else 0},
           "note": "the scan (1.2 MB) is shared by all hypotheses and is L2-resident by design; 8 scans are rotated between batches"}
    if px is not None:
        px.close()
    return out
