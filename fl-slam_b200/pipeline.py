"""The fused per-scan LiDAR evidence path: PointBudgetResample -> DeskewConstantTwist -> ray directions ->
BinSoftAssign -> ScanBinMomentMatch (+Kappa) -> MatrixFisherRotation -> PlanarTranslationEvidence -> 22-D evidence
(README.md:105-120 of the reference; call sites pkg/backend/pipeline.py:400,569,587-593).

Three launches per scan (or per whole scan stream): scan_moments (the only O(N) kernel), reduce_finalize, and the
chain step (map-derived statistics, pose evidence, measurement-noise IW update, map update).  Nothing synchronises with the host until the caller asks for certificates.  Everything runs in
libgcl_b200.so; without it this module cannot be imported.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi, constants
from .certificates import CertBundle, ExpectedEffect, InfluenceCert, SupportCert
from . import operators as ops

_lib = _abi.lib
PRECISIONS = {"fp64": _abi.GCL_F64, "fp32c": _abi.GCL_F32C}


_EMPTY_XI: Dict[int, np.ndarray] = {}   # (H, 0) placeholders: the message path carries no host-side twist array


@dataclass
class ResidentScan:
    """A raw scan resident in HBM, in one of the two layouts of include/gcl_b200.h."""
    points: torch.Tensor                 # (n,3) f64  |  (n,4) f32 {x,y,z,t - scan_start}
    timestamps: Optional[torch.Tensor]   # (n,) f64   |  None
    weights: torch.Tensor                # (n,) f64   |  (n,) f32
    n_raw: int
    layout: int
    scan_start_time: float
    scan_end_time: float
    row_offset: int = 0                  # raw index of row 0 of these tensors (point sharding: a rank holds a slice)

    @property
    def bytes_per_point(self) -> int:
        return 40 if self.layout == _abi.GCL_LAYOUT_REF_F64 else 20


@dataclass
class LidarScanEvidence:
    """Everything the chain produces for one scan under H hypotheses.  Tensors stay on the device."""
    n_bins: int
    n_hyp: int
    stats: torch.Tensor        # (H, stats record)
    evidence: torch.Tensor     # (H, GCL_EVID_DOUBLES)
    stats_host: Optional[np.ndarray] = None
    evidence_host: Optional[np.ndarray] = None
    chart_id: str = constants.GC_CHART_ID
    anchor_id: str = "initial"
    n_input: int = 0
    n_points_cap: int = 0
    ess_imu: float = 0.0

    # -- typed views ------------------------------------------------------------------------------------------
    def scan_bins(self, h: int = 0) -> ops.ScanBinStats:
        return ops.stats_views(self.stats[h], self.n_bins)

    def L_lidar(self, h: int = 0) -> torch.Tensor:
        o = _abi.EV["L22"]
        return self.evidence[h, o:o + 484].view(22, 22)

    def h_lidar(self, h: int = 0) -> torch.Tensor:
        o = _abi.EV["H22"]
        return self.evidence[h, o:o + 22]

    def _sync(self):
        if self.stats_host is None:
            self.stats_host = self.stats.detach().cpu().numpy()
            self.evidence_host = self.evidence.detach().cpu().numpy()

    def scalars(self, h: int = 0) -> Dict[str, float]:
        self._sync()
        tail = self.stats_host[h, 38 * self.n_bins:]
        return {k: float(tail[i]) for k, i in _abi.SS.items()}

    # -- certificates in pipeline order --------------------------------------------------------------------------
    def certificates(self, h: int = 0) -> List[Tuple[str, Any, CertBundle, ExpectedEffect]]:
        """[(operator name, result, CertBundle, ExpectedEffect)] in the order the reference pipeline emits them."""
        self._sync()
        s = self.scalars(h)
        eps = constants.GC_EPS_MASS
        out = []
        ess = s["PB_ESS"]
        pb_cert = CertBundle.create_approx(
            chart_id=self.chart_id, anchor_id=self.anchor_id, triggers=["PointBudgetResample"],
            support=SupportCert(ess_total=ess, support_frac=float(min(1.0, self.n_points_cap / (self.n_input + eps)))),
            influence=InfluenceCert.identity().with_overrides(mass_epsilon_ratio=eps / (s["TOTAL_MASS_IN"] + eps)))
        out.append(("PointBudgetResample", {"n_input": self.n_input, "n_output": int(s["N_SELECTED"]), "stride": int(s["STRIDE"]),
                                            "total_mass_in": s["TOTAL_MASS_IN"], "total_mass_out": s["TOTAL_MASS_IN"]},
                    pb_cert, ExpectedEffect("predicted_ess", ess)))
        dk_cert = CertBundle.create_exact(chart_id=self.chart_id, anchor_id=self.anchor_id,
                                          support=SupportCert(ess_total=float(self.ess_imu), support_frac=s["DESKEW_RETAINED"]),
                                          influence=InfluenceCert.identity())
        out.append(("DeskewConstantTwist", {"sum_w_out": s["SUM_W_OUT"]}, dk_cert,
                    ExpectedEffect("deskew_variance_reduction_proxy", 0.0)))
        sa_cert = CertBundle.create_exact(chart_id=self.chart_id, anchor_id=self.anchor_id,
                                          support=SupportCert(ess_total=float(math.exp(s["AVG_ENTROPY"])), support_frac=s["MAX_RESP"]))
        out.append(("BinSoftAssign", {"avg_entropy_valid_rows": s["ENTROPY_VALID_ROWS"], "max_resp_valid_rows": s["MAX_RESP_VALID_ROWS"]},
                    sa_cert, ExpectedEffect("predicted_assignment_entropy", s["AVG_ENTROPY"])))
        mm_cert, mm_eff = ops._moment_match_cert(self.stats_host[h, 38 * self.n_bins:], self.chart_id, self.anchor_id)
        out.append(("ScanBinMomentMatch", self.scan_bins(h), mm_cert, mm_eff))
        mf, mf_cert, mf_eff = ops.mf_result_from_record(self.evidence[h], self.evidence_host[h], self.chart_id, self.anchor_id)
        out.append(("MatrixFisherRotationEvidence", mf, mf_cert, mf_eff))
        pt, pt_cert, pt_eff = ops.pt_result_from_record(self.evidence[h], self.evidence_host[h], self.chart_id, self.anchor_id)
        out.append(("PlanarTranslationEvidence", pt, pt_cert, pt_eff))
        return out


class LidarEvidencePipeline:
    """Owns the atlas and the work buffers for one (n_bins, tau, precision) configuration on one GPU."""

    def __init__(self, n_bins: int = constants.GC_B_BINS, tau: float = None, precision: str = "fp64",
                 n_points_cap: int = constants.GC_N_POINTS_CAP, lidar_origin: Sequence[float] = (0.0, 0.0, 0.0),
                 eps_psd: float = constants.GC_EPS_PSD, eps_mass: float = constants.GC_EPS_MASS, bin_directions=None,
                 device: Optional[int] = None):
        if tau is None:
            raise ValueError("tau is required: the reference ships no GC_TAU_SOFT_ASSIGN (SURVEY.md 0.4)")
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        if not torch.cuda.is_available():
            raise _abi.GclError("no CUDA device: fl_slam_b200 has no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.n_bins, self.tau, self.precision = int(n_bins), float(tau), precision
        self.n_points_cap = int(n_points_cap)
        self.origin = np.asarray(lidar_origin, dtype=np.float64).reshape(3)
        self.eps_psd, self.eps_mass = float(eps_psd), float(eps_mass)
        with torch.cuda.device(self.device):
            self.handle = _abi.handle(self.device_index)
            if bin_directions is None:
                self.bin_dirs = ops.create_fibonacci_atlas(self.n_bins).dirs
            else:
                self.bin_dirs = ops._dev(bin_directions).reshape(self.n_bins, 3)
        self.rec = int(_lib.gcl_partial_record_doubles(self.n_bins))
        self.srec = int(_lib.gcl_stats_record_doubles(self.n_bins))
        self._bufs: Dict[Tuple[int, int], Dict[str, torch.Tensor]] = {}
        self._origin_c = _abi.dvec(self.origin)
        self._ptr_cache: Dict[int, Tuple[np.ndarray, int]] = {}
        self._inflight: Dict[int, tuple] = {}
        self._msg_layouts: Dict[tuple, tuple] = {}
        self._map_checked: Optional[tuple] = None
        self._ov: Optional[Dict[str, Any]] = None          # streams / events of run_stream(overlap=True)

    # ------------------------------------------------------------------------------------------------------------
    def manifest_entries(self, n_gpus: int = 1, shard_mode: str = "none") -> Dict[str, Any]:
        from .manifest import runtime_manifest_entries
        return runtime_manifest_entries(self, n_gpus=n_gpus, shard_mode=shard_mode)

    def upload_scan(self, points, timestamps, weights, scan_start_time: float, scan_end_time: float,
                    packed: Optional[bool] = None, row_offset: int = 0) -> ResidentScan:
        """Host (or device) arrays -> a scan resident in HBM.  packed=True converts to the 20 B/point layout on the
        device (fp32c mode only); default: packed iff precision == 'fp32c'."""
        packed = (self.precision == "fp32c") if packed is None else bool(packed)
        if packed and self.precision != "fp32c":
            raise ValueError("the packed fp32 layout needs precision='fp32c'")
        with torch.cuda.device(self.device):
            pts, ts, w = ops._dev(points).reshape(-1, 3), ops._dev(timestamps).reshape(-1), ops._dev(weights).reshape(-1)
            n = pts.shape[0]
            if not packed:
                return ResidentScan(pts, ts, w, n, _abi.GCL_LAYOUT_REF_F64, float(scan_start_time), float(scan_end_time), row_offset)
            xyzt = torch.empty((n, 4), dtype=torch.float32, device=self.device)
            wf = torch.empty((n,), dtype=torch.float32, device=self.device)
            _abi.check(_lib.gcl_pack_points_f32(self.handle, ops._stream(), ops._ptr(pts), ops._ptr(ts), ops._ptr(w), n,
                                                float(scan_start_time), ops._ptr(xyzt), ops._ptr(wf)), self.handle)
            return ResidentScan(xyzt, None, wf, n, _abi.GCL_LAYOUT_PACKED_F32, float(scan_start_time), float(scan_end_time), row_offset)

    # ------------------------------------------------------------------------------------------------------------
    def make_job(self, scan: ResidentScan, xi_body, n_raw_total: Optional[int] = None, sel_range: Tuple[int, int] = (0, -1),
                 point_lambda: Optional[torch.Tensor] = None, n_points_cap: Optional[int] = None) -> _abi.GclJob:
        """A gcl_job for `scan`.  When the scan tensor holds only raw rows [row_offset, row_offset + len) of a larger
        cloud (point sharding) the job carries virtual base pointers so that global raw indices address it."""
        j = _abi.GclJob()
        rowb = 24 if scan.layout == _abi.GCL_LAYOUT_REF_F64 else 16
        wb = 8 if scan.layout == _abi.GCL_LAYOUT_REF_F64 else 4
        off = int(scan.row_offset)
        j.points = scan.points.data_ptr() - off * rowb
        j.timestamps = (scan.timestamps.data_ptr() - off * 8) if scan.timestamps is not None else None
        j.weights = scan.weights.data_ptr() - off * wb
        j.point_lambda = None if point_lambda is None else point_lambda.data_ptr()
        j.n_raw = int(scan.n_raw if n_raw_total is None else n_raw_total)
        j.n_points_cap = int(self.n_points_cap if n_points_cap is None else n_points_cap)
        j.sel_begin, j.sel_end = int(sel_range[0]), int(sel_range[1])
        j.scan_start_time, j.scan_end_time = scan.scan_start_time, scan.scan_end_time
        xi = np.asarray(xi_body, dtype=np.float64).reshape(6)
        for k in range(6):
            j.xi_body[k] = float(xi[k])
        return j

    def jobs_to_device(self, jobs: Sequence[_abi.GclJob]) -> torch.Tensor:
        arr = (_abi.GclJob * len(jobs))(*jobs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        return host.to(self.device, non_blocking=False)

    @staticmethod
    def selected_rows(n_raw: int, cap: int) -> Tuple[int, int]:
        """(stride, n_selected) of point_budget.py:160,:70-71."""
        stride = max(1, int(math.ceil(n_raw / cap))) if n_raw > 0 else 1
        return stride, (n_raw + stride - 1) // stride

    def _buffers(self, n_jobs: int, ctas: int, slot: int = 0) -> Dict[str, torch.Tensor]:
        key = (n_jobs, ctas, slot)
        b = self._bufs.get(key)
        if b is None:
            z = lambda *s: torch.empty(s, dtype=torch.float64, device=self.device)
            b = {"partials": z(n_jobs, ctas, self.rec), "reduced": z(n_jobs, self.rec), "stats": z(n_jobs, self.srec),
                 "evidence": z(n_jobs, _abi.GCL_EVID_DOUBLES), "map_derived": z(max(int(_lib.gcl_scan_chain_workspace_doubles(n_jobs, 1, self.n_bins)),
                                      int(_lib.gcl_scan_chain_workspace_doubles(1, n_jobs, self.n_bins))))}
            self._bufs[key] = b
        return b

    def plan_ctas(self, n_jobs: int, max_sel_rows: int) -> int:
        return int(_lib.gcl_plan_ctas_per_job(self.handle, int(n_jobs), int(max_sel_rows), self.n_bins, PRECISIONS[self.precision]))

    # ---- the launches (all asynchronous on the current stream) ------------------------------------------------------
    def launch_moments(self, jobs_dev: Optional[torch.Tensor], job: Optional[_abi.GclJob], n_jobs: int, ctas: int, layout: int,
                       partials: torch.Tensor) -> None:
        _abi.check(_lib.gcl_scan_moments(self.handle, ops._stream(), None if jobs_dev is None else C.c_void_p(jobs_dev.data_ptr()),
                                         None if job is None else C.byref(job), n_jobs, ctas, layout, PRECISIONS[self.precision],
                                         ops._ptr(self.bin_dirs), self.n_bins, self.tau, self._origin_c, self.eps_mass,
                                         ops._ptr(partials)), self.handle)

    def launch_finalize(self, parts: torch.Tensor, n_parts: int, part_stride: int, job_stride: int,
                        jobs_dev: Optional[torch.Tensor], job: Optional[_abi.GclJob], n_jobs: int, finalize: bool,
                        reduced: torch.Tensor, stats: Optional[torch.Tensor]) -> None:
        _abi.check(_lib.gcl_reduce_finalize(self.handle, ops._stream(), ops._ptr(parts), n_parts, part_stride, job_stride,
                                            None if jobs_dev is None else C.c_void_p(jobs_dev.data_ptr()),
                                            None if job is None else C.byref(job), n_jobs, self.n_bins, ops._ptr(self.bin_dirs),
                                            self.tau, self._origin_c, self.eps_psd, self.eps_mass, 1 if finalize else 0,
                                            ops._ptr(reduced), ops._ptr(stats)), self.handle)

    def launch_map_derived(self, m: ops.MapBinStats, out: torch.Tensor) -> None:
        _abi.check(_lib.gcl_map_derived_stats(self.handle, ops._stream(), ops._ptr(m.S_dir), ops._ptr(m.N_dir), ops._ptr(m.N_pos),
                                              ops._ptr(m.sum_p), ops._ptr(m.sum_ppT), self.n_bins, self.eps_mass, self.eps_psd,
                                              ops._ptr(out)), self.handle)

    def launch_evidence(self, stats: torch.Tensor, stats_stride: int, m: ops.MapBinStats, map_derived: torch.Tensor,
                        poses: torch.Tensor, n_jobs: int, evidence: torch.Tensor) -> None:
        _abi.check(_lib.gcl_pose_evidence(self.handle, ops._stream(), ops._ptr(stats), stats_stride, ops._ptr(m.S_dir),
                                          ops._ptr(m.S_dir_scatter), ops._ptr(m.N_dir), ops._ptr(m.N_pos), ops._ptr(map_derived),
                                          ops._ptr(poses), None, n_jobs, self.n_bins, self.eps_psd, self.eps_mass, 1, 1,
                                          ops._ptr(evidence)), self.handle)

    def launch_chain(self, stats: torch.Tensor, n_scans: int, n_hyp: int, m: ops.MapBinStats, map_derived: torch.Tensor,
                     poses_pred: torch.Tensor, poses_update: Optional[torch.Tensor], evidence: torch.Tensor,
                     iw_state: Optional[torch.Tensor] = None, use_sigma_meas: bool = False, forgetting: float = 0.99, phase: int = 0) -> None:
        """The sequential part of a stream in one launch (gcl_scan_chain): map + IW state stay on the device.  ``phase`` 1 / 2:
        the map-dependent and the IW-dependent half (gcl_scan_chain_phase)."""
        _abi.check(_lib.gcl_scan_chain_phase(self.handle, ops._stream(), int(phase), ops._ptr(stats), self.srec, n_scans, n_hyp, self.n_bins,
                                       ops._ptr(m.S_dir), ops._ptr(m.S_dir_scatter), ops._ptr(m.N_dir), ops._ptr(m.N_pos),
                                       ops._ptr(m.sum_p), ops._ptr(m.sum_ppT), ops._ptr(map_derived), ops._ptr(poses_pred),
                                       ops._ptr(poses_update), float(forgetting), ops._ptr(iw_state), 1 if use_sigma_meas else 0,
                                       constants.GC_IW_RHO_MEAS_LIDAR, constants.GC_IW_NU_WEAK_ADD, constants.GC_IW_NU_MAX,
                                       self.eps_psd, self.eps_mass, ops._ptr(evidence)), self.handle)

    def create_lidar_iw_state(self, sigma_meas: float = constants.GC_LIDAR_SIGMA_MEAS) -> torch.Tensor:
        """[nu, Psi(3x3)] of the LiDAR measurement-noise block: nu = p + 1 + nu_extra, Psi = sigma I nu_extra
        (pkg/backend/structures/measurement_noise_iw_jax.py:36-68)."""
        v = np.concatenate([[3.0 + 1.0 + constants.GC_IW_NU_WEAK_ADD], (sigma_meas * constants.GC_IW_NU_WEAK_ADD * np.eye(3)).reshape(-1)])
        return torch.from_numpy(v).to(self.device)

    # ---- a stream of scans resident in HBM: three launches for the whole stream -------------------------------------------
    def run_stream(self, scans: Sequence[ResidentScan], xi_bodies, poses_pred: torch.Tensor, poses_update: Optional[torch.Tensor],
                   map_stats: ops.MapBinStats, iw_state: Optional[torch.Tensor] = None, use_sigma_meas: bool = False,
                   forgetting: float = 0.99, jobs_dev: Optional[torch.Tensor] = None, ctas: Optional[int] = None,
                   overlap: bool = False) -> LidarScanEvidence:
        """scan_moments and reduce_finalize run for ALL scans at once (they do not depend on the map); one chain launch
        then walks the scans in order.  Returns per-scan stats / evidence records (n_hyp = number of scans here).

        ``overlap=True`` pipelines consecutive batches: the moment pass of this batch runs on an internal stream and only the
        chain step (map + IW state carried from the previous batch) is ordered after the previous batch's chain, on a second,
        high-priority internal stream -- the sequential chain keeps one SM busy while the next batch's moment pass fills the other
        147.  Inputs must be ready on the current stream when the call is made; results alternate between two buffer sets and are
        ordered on the current stream by ``join_stream()`` (they stay valid until the call after next).  Inputs produced on the
        current stream before the call are ordered before the batch (the internal streams wait for it at every call)."""
        S = len(scans)
        xi = np.asarray(xi_bodies, dtype=np.float64).reshape(S, 6)
        max_sel = max(self.selected_rows(s.n_raw, self.n_points_cap)[1] for s in scans)
        ctas = self.plan_ctas(S, max_sel) if ctas is None else ctas
        if jobs_dev is None:
            jobs_dev = self.jobs_to_device([self.make_job(scans[k], xi[k]) for k in range(S)])
        if not overlap:
            b = self._buffers(S, ctas)
            self.launch_moments(jobs_dev, None, S, ctas, scans[0].layout, b["partials"])
            self.launch_finalize(b["partials"], ctas, self.rec, self.rec * ctas, jobs_dev, None, S, True, b["reduced"], b["stats"])
            self.launch_chain(b["stats"], S, 1, map_stats, b["map_derived"], poses_pred, poses_update, b["evidence"], iw_state,
                              use_sigma_meas, forgetting)
        else:
            ov = self._ov
            if ov is None:
                ov = self._ov = {"s_mom": torch.cuda.Stream(device=self.device, priority=0), "s_chain": torch.cuda.Stream(device=self.device, priority=-1),
                                 "free": [None, None], "turn": 0, "last": None, "keep": [None, None]}
            # EVERY call: what the caller enqueued on the current stream so far (uploads of scans, poses, map edits, the jobs
            # array built above) comes before this batch on both internal streams
            cur = torch.cuda.current_stream(self.device)
            ov["s_mom"].wait_stream(cur); ov["s_chain"].wait_stream(cur)
            i = ov["turn"] & 1
            ov["turn"] += 1
            b = self._buffers(S, ctas, slot=i)
            # caller tensors read by the internal streams stay referenced until this buffer set comes round again (two calls
            # later, after its chain has been waited for): the caching allocator cannot hand their memory out in between
            ov["keep"][i] = (jobs_dev, poses_pred, poses_update, iw_state, list(scans))
            for t in (jobs_dev, poses_pred, poses_update, iw_state):
                if t is not None:
                    t.record_stream(ov["s_mom"]); t.record_stream(ov["s_chain"])
            if ov["free"][i] is not None:
                ov["s_mom"].wait_event(ov["free"][i])                            # this buffer set was last read by the chain two calls ago
            with torch.cuda.stream(ov["s_mom"]):
                self.launch_moments(jobs_dev, None, S, ctas, scans[0].layout, b["partials"])
                self.launch_finalize(b["partials"], ctas, self.rec, self.rec * ctas, jobs_dev, None, S, True, b["reduced"], b["stats"])
                # the half of the chain that depends only on the map (left by the previous batch's phase 1, same stream)
                self.launch_chain(b["stats"], S, 1, map_stats, b["map_derived"], poses_pred, poses_update, b["evidence"], iw_state,
                                  use_sigma_meas, forgetting, phase=1)
                ready = ov["s_mom"].record_event()
            ov["s_chain"].wait_event(ready)
            with torch.cuda.stream(ov["s_chain"]):
                # the half that depends on the IW state (left by the previous batch's phase 2, same stream)
                self.launch_chain(b["stats"], S, 1, map_stats, b["map_derived"], poses_pred, poses_update, b["evidence"], iw_state,
                                  use_sigma_meas, forgetting, phase=2)
                ov["free"][i] = ov["last"] = ov["s_chain"].record_event()
        return LidarScanEvidence(n_bins=self.n_bins, n_hyp=S, stats=b["stats"], evidence=b["evidence"], n_input=scans[0].n_raw,
                                 n_points_cap=self.n_points_cap)

    def join_stream(self) -> None:
        """Order the current stream after every batch submitted with ``run_stream(overlap=True)``."""
        if self._ov is not None and self._ov["last"] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._ov["last"])

    # ---- one scan, H hypotheses, data already in HBM -------------------------------------------------------------------
    def run_resident(self, scan: ResidentScan, xi_bodies, poses_dev: torch.Tensor, map_stats: ops.MapBinStats,
                     point_lambda: Optional[torch.Tensor] = None, jobs_dev: Optional[torch.Tensor] = None,
                     ctas: Optional[int] = None, pose_update: Optional[torch.Tensor] = None,
                     iw_state: Optional[torch.Tensor] = None, use_sigma_meas: bool = False, forgetting: float = 0.99
                     ) -> LidarScanEvidence:
        """Device-resident pass, three launches, no host synchronisation.  xi_bodies: (H,6) host array; poses_dev: (H,6)
        device.  pose_update (1,6): if given, the map statistics are updated in place after the evidence (hypothesis 0);
        iw_state: LiDAR measurement-noise IW block updated in place."""
        xi = np.asarray(xi_bodies, dtype=np.float64).reshape(-1, 6)
        H = xi.shape[0]
        _, n_sel = self.selected_rows(scan.n_raw, self.n_points_cap)
        ctas = self.plan_ctas(H, n_sel) if ctas is None else ctas
        b = self._buffers(H, ctas)
        job = None
        if H == 1 and jobs_dev is None:
            job = self.make_job(scan, xi[0], point_lambda=point_lambda)
        elif jobs_dev is None:
            jobs_dev = self.jobs_to_device([self.make_job(scan, xi[h], point_lambda=point_lambda) for h in range(H)])
        self.launch_moments(jobs_dev, job, H, ctas, scan.layout, b["partials"])
        self.launch_finalize(b["partials"], ctas, self.rec, self.rec * ctas, jobs_dev, job, H, True, b["reduced"], b["stats"])
        self.launch_chain(b["stats"], 1, H, map_stats, b["map_derived"], poses_dev, pose_update, b["evidence"], iw_state,
                          use_sigma_meas, forgetting)
        return LidarScanEvidence(n_bins=self.n_bins, n_hyp=H, stats=b["stats"], evidence=b["evidence"], n_input=scan.n_raw,
                                 n_points_cap=self.n_points_cap)

    # ---- the reference-facing calls: host arrays in, certificates out -----------------------------------------------------
    _MAP_KEYS = ("S_dir", "S_dir_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT")

    def _map_shapes(self):
        B = self.n_bins
        return (("S_dir", (B, 3)), ("S_dir_scatter", (B, 3, 3)), ("N_dir", (B,)), ("N_pos", (B,)), ("sum_p", (B, 3)), ("sum_ppT", (B, 3, 3)))

    @staticmethod
    def _f64(a, shape):
        """float64 C-contiguous view/copy of `a` with `shape` (no copy, no new object when `a` already is one)."""
        if type(a) is np.ndarray and a.dtype == np.float64 and a.flags.c_contiguous:
            return a if a.shape == shape else a.reshape(shape)
        return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(shape))

    def _addr(self, a: Optional[np.ndarray]) -> Optional[int]:
        """Address of a NumPy array; arrays that come back call after call (map statistics, IW state) are looked up."""
        if a is None:
            return None
        hit = self._ptr_cache.get(id(a))
        if hit is not None and hit[0] is a:
            return hit[1]
        addr = a.ctypes.data
        if len(self._ptr_cache) > 512:   # small arrays only (map statistics, IW state, IMU buffers, extrinsics)
            self._ptr_cache.clear()
        self._ptr_cache[id(a)] = (a, addr)   # holding `a` keeps id(a) from being reused by another object
        return addr

    def register_pinned(self, *arrays: np.ndarray) -> None:
        """Vouch for page-locked arrays that carry scans call after call (gcl_host_register): the calls that read them in place
        then skip their per-call driver query.  Raises ValueError for pageable memory."""
        for a in arrays:
            _abi.check(_lib.gcl_host_register(self.handle, a.ctypes.data, a.nbytes), self.handle)

    def unregister_pinned(self, *arrays: np.ndarray) -> None:
        for a in arrays:
            _abi.check(_lib.gcl_host_unregister(self.handle, a.ctypes.data), self.handle)

    def submit_scan_host(self, slot: int, points, timestamps, weights, scan_start_time: float, scan_end_time: float, xi_body) -> int:
        """Stage 1 of the host call (gcl_scan_submit_host): asynchronous upload + moment pass of one scan under H twists.
        Returns H.  Page-locked inputs are read in place: keep them unchanged until ``evidence_host(slot, ...)`` returns."""
        c = self._c64
        pts, ts, w = c(points, 0, 3), c(timestamps, 0, 1), c(weights, 0, 1)   # only sizes and addresses matter below
        n = ts.size
        if pts.size != 3 * n or w.size != n:
            raise ValueError("points, timestamps and weights must have the same length")
        xi = c(xi_body, 0, 6)
        if xi.ndim != 2:
            xi = xi.reshape(-1, 6)
        bins_host = getattr(self, "_bins_host", None)
        if bins_host is None:
            bins_host = self._bins_host = np.ascontiguousarray(self.bin_dirs.detach().cpu().numpy())
        self._inflight[slot] = (pts, ts, w, xi, n)           # keep the arrays alive while the device reads them
        _abi.check(_lib.gcl_scan_submit_host(self.handle, int(slot), pts.ctypes.data, ts.ctypes.data, w.ctypes.data, n, self.n_points_cap,
                                             float(scan_start_time), float(scan_end_time), self._addr(xi), xi.shape[0], self._addr(self.origin),
                                             self._addr(bins_host), self.n_bins, self.tau, PRECISIONS[self.precision], self.eps_psd,
                                             self.eps_mass), self.handle)
        return xi.shape[0]

    def submit_scan_msg_host(self, slot: int, msg=None, *, data=None, fields=None, point_step: Optional[int] = None,
                             n_points: Optional[int] = None, header_stamp_sec: float = 0.0, R_base_lidar=None, t_base_lidar=None,
                             xi_body=None, imu: Optional[Dict[str, Any]] = None) -> int:
        """Stage 1 from the MESSAGE (gcl_scan_submit_msg_host): the raw PointCloud2 bytes are uploaded as they are and parsed,
        windowed, transformed to the base frame and resampled on the device; the scan time bounds
        (pkg/backend/backend_node.py:1776-1786) are computed there too.  The deskew twist is either ``xi_body`` (H,6) or comes
        from ``imu`` = dict(stamps, gyro, accel, sigma, rotvec_start_WB (H,3), gyro_bias (H,3), accel_bias (H,3), gravity_W,
        deskew_rotation_only=False), preintegrated on the device over those bounds (pkg/backend/pipeline.py:437-483).
        ``msg`` is a sensor_msgs/PointCloud2-like object, or pass the pieces by keyword (``fields`` as (name, offset,
        datatype) triples).  A page-locked ``data`` array (np.uint8) is read in place.  Returns H."""
        if (xi_body is None) == (imu is None):
            raise ValueError("exactly one of xi_body and imu must be given")
        if msg is not None:
            data, point_step, n_points = msg.data, int(msg.point_step), int(msg.width) * int(msg.height)
            fields = [(f.name, int(f.offset), int(f.datatype)) for f in msg.fields]
            header_stamp_sec = msg.header.stamp.sec + msg.header.stamp.nanosec * 1e-9
        key = (tuple(fields), int(point_step))
        lay = self._msg_layouts.get(key)
        if lay is None:
            fmap = {f[0]: (int(f[1]), int(f[2])) for f in fields}
            missing = [k for k in ("x", "y", "z", "ring") if k not in fmap]
            if missing:
                raise RuntimeError(f"PointCloud2 (VLP-16 layout) missing required fields: {missing}. Present fields: {sorted(fmap)}")
            tname = "t" if "t" in fmap else ("time" if "time" in fmap else None)
            order = ["x", "y", "z", "ring"] + ([tname] if tname else [])
            off = np.array([fmap[k][0] for k in order] + ([] if tname else [-1]), dtype=np.int32)
            dt = np.array([fmap[k][1] for k in order] + ([] if tname else [0]), dtype=np.int32)
            if any(d < 1 or d > 8 for d in dt[:len(order)]):
                raise ValueError(f"Unsupported PointField datatype: {[int(d) for d in dt]}")
            lay = self._msg_layouts[key] = (off, dt, off.ctypes.data, dt.ctypes.data)
        off_p, dt_p = lay[2], lay[3]
        n = int(n_points)
        raw = data if (type(data) is np.ndarray and data.dtype == np.uint8 and data.flags.c_contiguous) else np.frombuffer(bytes(data), dtype=np.uint8)
        if raw.size < n * int(point_step):
            raise ValueError("message data is shorter than n_points * point_step")
        c, ad = self._c64, self._addr
        R = None if R_base_lidar is None else c(R_base_lidar, 9)
        t = None if t_base_lidar is None else c(t_base_lidar, 3)
        if xi_body is not None:
            xi = c(xi_body, 0, 6)
            H = xi.size // 6
            imu_args = (ad(xi), None, None, None, 0, 0.0, None, None, None, None, 0)
            keep = (xi,)
        else:
            st, gy, ac = c(imu["stamps"], 0, 1), c(imu["gyro"], 0, 3), c(imu["accel"], 0, 3)
            M = st.size
            if gy.size != 3 * M or ac.size != 3 * M:
                raise ValueError("imu stamps, gyro and accel must have the same length")
            rv, bg, ba = c(imu["rotvec_start_WB"], 0, 3), c(imu["gyro_bias"], 0, 3), c(imu["accel_bias"], 0, 3)
            H = rv.size // 3
            if bg.size != 3 * H or ba.size != 3 * H:
                raise ValueError("rotvec_start_WB, gyro_bias and accel_bias must have one row per hypothesis")
            grav = c(imu["gravity_W"], 3)
            imu_args = (None, ad(st), ad(gy), ad(ac), M, float(imu["sigma"]), ad(rv), ad(bg), ad(ba), ad(grav),
                        1 if imu.get("deskew_rotation_only", False) else 0)
            keep = (st, gy, ac, rv, bg, ba, grav)
        bins_host = getattr(self, "_bins_host", None)
        if bins_host is None:
            bins_host = self._bins_host = np.ascontiguousarray(self.bin_dirs.detach().cpu().numpy())
        ph = _EMPTY_XI.get(H)
        if ph is None:
            ph = _EMPTY_XI[H] = np.empty((H, 0))
        self._inflight[slot] = (raw, keep, None, ph, n)   # arrays stay alive while the copy is in flight
        _abi.check(_lib.gcl_scan_submit_msg_host(self.handle, int(slot), raw.ctypes.data, n, int(point_step), off_p, dt_p, float(header_stamp_sec),
                                                 ad(R), ad(t), self.n_points_cap, *imu_args, H, ad(self.origin), ad(bins_host),
                                                 self.n_bins, self.tau, PRECISIONS[self.precision], self.eps_psd, self.eps_mass), self.handle)
        return H

    @staticmethod
    def _c64(a, size: int, mult: int = 1):
        """``a`` as a float64 C-contiguous array whose size is ``size`` (or, with size 0, a multiple of ``mult``): no copy and no
        new object when it already is one (only its address is needed, so the shape itself does not matter)."""
        if not (type(a) is np.ndarray and a.dtype == np.float64 and a.flags.c_contiguous):
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if (size and a.size != size) or (not size and a.size % mult):
            raise ValueError(f"array of {a.size} elements where {'%d' % size if size else 'a multiple of %d' % mult} were expected")
        return a

    def msg_info_host(self, slot: int, n_hyp: int = 1) -> Dict[str, Any]:
        """What the device derived from the message submitted to ``slot`` (gcl_scan_msg_info_host): the scan time bounds, the
        time unit, the resample stride and, per hypothesis, the IMU preintegration record with the deskew twist."""
        out = np.empty(_abi.GCL_MSG_INFO_HEADER + n_hyp * ops.GCL_IMU_DOUBLES)
        _abi.check(_lib.gcl_scan_msg_info_host(self.handle, int(slot), out.ctypes.data, out.size), self.handle)
        recs = out[_abi.GCL_MSG_INFO_HEADER:].reshape(n_hyp, ops.GCL_IMU_DOUBLES)
        return {"scan_start_time": float(out[0]), "scan_end_time": float(out[1]), "t_min": float(out[2]), "t_max": float(out[3]),
                "time_in_nanoseconds": bool(out[4]), "n_points": int(out[5]), "stride": int(out[6]), "n_selected": int(out[7]),
                "imu_records": recs, "xi_body": recs[:, ops.IMU["XI_BODY"]:ops.IMU["XI_BODY"] + 6], "ess_imu": recs[:, ops.IMU["ESS"]]}

    def evidence_begin_host(self, slot: int, pose_pred, map_stats_host: Dict[str, np.ndarray], ess_imu: float = 0.0,
                            chart_id: str = constants.GC_CHART_ID, anchor_id: str = "initial", pose_update=None,
                            forgetting: float = 0.99, iw_state_host: Optional[np.ndarray] = None, use_sigma_meas: bool = False) -> None:
        """First half of stage 2 (gcl_scan_evidence_begin_host): the chain step of the scan submitted to ``slot`` and its one
        D2H are enqueued; nothing is waited for, so the next scan can be submitted while the GPU works.

        With ``pose_update`` (6,) the six arrays of ``map_stats_host`` (float64, C-contiguous) are updated IN PLACE by
        ``evidence_end_host`` (forgetting + this scan's bin sums in the world frame); ``iw_state_host`` (10,) likewise."""
        f = self._f64
        pts, ts, w, xi, n = self._inflight.pop(slot)
        H = xi.shape[0]
        poses = f(pose_pred, (-1, 6))
        if poses.shape[0] != H:
            raise ValueError("xi_body and pose_pred must have one row per hypothesis")
        arrs, ma = self._map_args(map_stats_host, iw_state_host, pose_update is not None)
        p = self._addr
        pu = None if pose_update is None else f(pose_update, (6,))
        _abi.check(_lib.gcl_scan_evidence_begin_host(self.handle, int(slot), poses.ctypes.data, ma[0], ma[1], ma[2], ma[3], ma[4], ma[5], p(pu),
                                                     float(forgetting), ma[6], 1 if use_sigma_meas else 0), self.handle)
        self._stage2 = (H, n, arrs, ma, ess_imu, chart_id, anchor_id, poses, pu)

    def _map_args(self, map_stats_host: Dict[str, np.ndarray], iw_state_host: Optional[np.ndarray], in_place: bool):
        """The six map arrays + the IW state as (arrays, addresses).  They usually are the SAME objects scan after scan (updated
        in place): their checks and addresses are remembered and re-used as long as every array is the very object that was
        checked."""
        f = self._f64
        mc = self._map_checked
        arrs = (map_stats_host["S_dir"], map_stats_host["S_dir_scatter"], map_stats_host["N_dir"], map_stats_host["N_pos"],
                map_stats_host["sum_p"], map_stats_host["sum_ppT"], iw_state_host)
        if mc is not None and mc[0][0] is arrs[0] and mc[0][1] is arrs[1] and mc[0][2] is arrs[2] and mc[0][3] is arrs[3] \
                and mc[0][4] is arrs[4] and mc[0][5] is arrs[5] and mc[0][6] is arrs[6]:
            ma = mc[1]
        else:
            shapes = self._map_shapes()
            conforming = all(isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == sh
                             for a, (k, sh) in zip(arrs, shapes))
            if in_place and not conforming:   # in/out: the caller's arrays are written, so they must be usable as they are
                raise ValueError("map_stats_host arrays must be C-contiguous float64 arrays of shapes "
                                 f"{dict(shapes)} for an in-place update")
            if iw_state_host is not None and not (isinstance(iw_state_host, np.ndarray) and iw_state_host.dtype == np.float64
                                                  and iw_state_host.shape == (10,) and iw_state_host.flags.c_contiguous):
                raise ValueError("iw_state_host must be a C-contiguous float64 array of shape (10,)")
            if conforming:
                ma = tuple(None if a is None else a.ctypes.data for a in arrs)
                self._map_checked = (arrs, ma)
            else:                                             # read-only use of non-conforming arrays: converted copies
                conv = tuple(f(a, sh) for a, (k, sh) in zip(arrs, shapes)) + (iw_state_host,)
                ma = tuple(None if a is None else a.ctypes.data for a in conv)
                arrs = conv
        return arrs, ma

    def evidence_end_host(self) -> LidarScanEvidence:
        """Second half of stage 2 (gcl_scan_evidence_end_host): wait for the D2H, write the map / IW arrays back, return."""
        H, n, _arrs, ma, ess_imu, chart_id, anchor_id, _poses, _pu = self._stage2
        self._stage2 = None
        stats = np.empty((H, self.srec), dtype=np.float64)
        evid = np.empty((H, _abi.GCL_EVID_DOUBLES), dtype=np.float64)
        _abi.check(_lib.gcl_scan_evidence_end_host(self.handle, ma[0], ma[1], ma[2], ma[3], ma[4], ma[5], ma[6], stats.ctypes.data,
                                                   evid.ctypes.data), self.handle)
        return LidarScanEvidence(n_bins=self.n_bins, n_hyp=H, stats=torch.from_numpy(stats), evidence=torch.from_numpy(evid),
                                 stats_host=stats, evidence_host=evid, chart_id=chart_id, anchor_id=anchor_id, n_input=n,
                                 n_points_cap=self.n_points_cap, ess_imu=ess_imu)

    def evidence_host(self, slot: int, pose_pred, map_stats_host: Dict[str, np.ndarray], **kw) -> LidarScanEvidence:
        """Stage 2 of the host call (= gcl_scan_evidence_host): chain step for the scan submitted to ``slot``, one D2H, sync."""
        self.evidence_begin_host(slot, pose_pred, map_stats_host, **kw)
        return self.evidence_end_host()

    def process_scan_host(self, points, timestamps, weights, scan_start_time: float, scan_end_time: float, xi_body, pose_pred,
                          map_stats_host: Dict[str, np.ndarray], ess_imu: float = 0.0, chart_id: str = constants.GC_CHART_ID,
                          anchor_id: str = "initial", pose_update=None, forgetting: float = 0.99,
                          iw_state_host: Optional[np.ndarray] = None, use_sigma_meas: bool = False) -> LidarScanEvidence:
        """The blocking reference-facing call (= gcl_lidar_scan_evidence_host): stage 1 and stage 2 back to back on slot 0."""
        xi = np.asarray(xi_body, dtype=np.float64).reshape(-1, 6)
        poses = np.asarray(pose_pred, dtype=np.float64).reshape(-1, 6)
        if poses.shape[0] != xi.shape[0]:
            raise ValueError("xi_body and pose_pred must have one row per hypothesis")
        c = self._c64
        pts, ts, w = c(points, 0, 3), c(timestamps, 0, 1), c(weights, 0, 1)
        n = ts.size
        if pts.size != 3 * n or w.size != n:
            raise ValueError("points, timestamps and weights must have the same length")
        xi, poses = np.ascontiguousarray(xi), np.ascontiguousarray(poses)
        H = xi.shape[0]
        arrs, ma = self._map_args(map_stats_host, iw_state_host, pose_update is not None)
        pu = None if pose_update is None else self._f64(pose_update, (6,))
        bins_host = getattr(self, "_bins_host", None)
        if bins_host is None:
            bins_host = self._bins_host = np.ascontiguousarray(self.bin_dirs.detach().cpu().numpy())
        stats = np.empty((H, self.srec), dtype=np.float64)
        evid = np.empty((H, _abi.GCL_EVID_DOUBLES), dtype=np.float64)
        ad = self._addr
        _abi.check(_lib.gcl_lidar_scan_evidence_host(
            self.handle, pts.ctypes.data, ts.ctypes.data, w.ctypes.data, n, self.n_points_cap, float(scan_start_time), float(scan_end_time),
            xi.ctypes.data, poses.ctypes.data, H, ad(self.origin), ad(bins_host), self.n_bins, self.tau, PRECISIONS[self.precision],
            ma[0], ma[1], ma[2], ma[3], ma[4], ma[5], ad(pu), float(forgetting), ma[6], 1 if use_sigma_meas else 0, self.eps_psd,
            self.eps_mass, stats.ctypes.data, evid.ctypes.data), self.handle)
        return LidarScanEvidence(n_bins=self.n_bins, n_hyp=H, stats=torch.from_numpy(stats), evidence=torch.from_numpy(evid),
                                 stats_host=stats, evidence_host=evid, chart_id=chart_id, anchor_id=anchor_id, n_input=n,
                                 n_points_cap=self.n_points_cap, ess_imu=ess_imu)


def lidar_scan_evidence(points, timestamps, weights, *, scan_start_time: float, scan_end_time: float, xi_body, belief_pred,
                        map_stats, lidar_origin, tau: float, n_bins: int = constants.GC_B_BINS,
                        n_points_cap: int = constants.GC_N_POINTS_CAP, precision: str = "fp64", ess_imu: float = 0.0,
                        pipeline: Optional[LidarEvidencePipeline] = None) -> LidarScanEvidence:
    """Fused drop-in for the operator sequence of the bin-era pipeline (README.md:110-115): one call, all certificates
    (``.certificates()``) in pipeline order, 22-D evidence in ``.L_lidar()`` / ``.h_lidar()``."""
    pl = pipeline or LidarEvidencePipeline(n_bins=n_bins, tau=tau, precision=precision, n_points_cap=n_points_cap,
                                           lidar_origin=lidar_origin)
    beliefs = belief_pred if isinstance(belief_pred, (list, tuple)) else [belief_pred]
    poses = np.stack([ops._pose6(b, constants.GC_EPS_LIFT) for b in beliefs])
    xi = np.asarray(xi_body, dtype=np.float64).reshape(-1, 6)
    if xi.shape[0] == 1 and len(beliefs) > 1:
        xi = np.repeat(xi, len(beliefs), axis=0)
    if isinstance(map_stats, ops.MapBinStats):
        msh = {k: getattr(map_stats, k).detach().cpu().numpy() for k in ("S_dir", "S_dir_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT")}
    else:
        msh = map_stats
    return pl.process_scan_host(points, timestamps, weights, scan_start_time, scan_end_time, xi, poses, msh, ess_imu=ess_imu,
                                chart_id=beliefs[0].chart_id, anchor_id=beliefs[0].anchor_id)
