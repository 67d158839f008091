"""Drop-in operators with the reference's names, arguments and return triples, backed by libgcl_b200.so.

Every function mirrors one reference operator (file:line in its docstring; pkg/ =
fl_ws/src/fl_slam_poc/fl_slam_poc/) and returns ``(result, CertBundle, ExpectedEffect)`` as
docs/OPERATOR_CONTRACTS.md:3 requires.  Arrays may be NumPy arrays or torch tensors (host or device); outputs are
float64 CUDA torch tensors owned by the caller.  Each operator performs ONE device-to-host copy of a small scalar
block to build its certificate (the reference does one sync per scalar, e.g. point_budget.py:185-199).

There is no CPU path: the CUDA library is required at import time (_abi.py).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Any, Optional, Tuple

import numpy as np
import torch

from . import _abi, constants
from .certificates import (CertBundle, ConditioningCert, ExpectedEffect, InfluenceCert, MismatchCert, SupportCert)

_lib = _abi.lib


# ----------------------------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------------------------
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _abi.GclError("no CUDA device: fl_slam_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _dev(x: Any, dtype=torch.float64) -> torch.Tensor:
    """To a contiguous tensor of `dtype` on the current CUDA device (inputs are never written to)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype={torch.float64: np.float64, torch.uint8: np.uint8,
                                                                       torch.float32: np.float32}[dtype])))
    return t.to(device=_device(), dtype=dtype).contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _h():
    return _abi.handle(torch.cuda.current_device())


def _empty(*shape, dtype=torch.float64) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=_device())


def _host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


@dataclass
class PosePrediction:
    """Minimal stand-in for the part of BeliefGaussianInfo the evidence operators read
    (pkg/common/belief.py:408-434): a 6-vector world pose [t, rotvec] plus chart / anchor ids."""
    pose: Any
    chart_id: str = constants.GC_CHART_ID
    anchor_id: str = "initial"

    def mean_world_pose(self, eps_lift: float = constants.GC_EPS_LIFT):
        return np.asarray(self.pose, dtype=np.float64).reshape(6)


def _pose6(belief_pred: Any, eps_lift: float) -> np.ndarray:
    p = belief_pred.mean_world_pose(eps_lift=eps_lift)
    if isinstance(p, torch.Tensor):
        p = p.detach().cpu().numpy()
    return np.asarray(p, dtype=np.float64).reshape(6)


# ----------------------------------------------------------------------------------------------------------------
# PointBudgetResample
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class PointBudgetResult:
    """pkg/backend/operators/point_budget.py:31-42."""
    points: torch.Tensor
    timestamps: torch.Tensor
    weights: torch.Tensor
    ring: torch.Tensor
    tag: torch.Tensor
    n_input: int
    n_output: int
    total_mass_in: float
    total_mass_out: float


def point_budget_resample(points, timestamps, weights, ring=None, tag=None, n_points_cap: int = constants.GC_N_POINTS_CAP,
                          chart_id: str = constants.GC_CHART_ID, anchor_id: str = "initial"
                          ) -> Tuple[PointBudgetResult, CertBundle, ExpectedEffect]:
    """pkg/backend/operators/point_budget.py:117-221 (deterministic stride subsample, mass-preserving rescale,
    zero-padded cap-sized outputs)."""
    pts, ts, w = _dev(points).reshape(-1, 3), _dev(timestamps).reshape(-1), _dev(weights).reshape(-1)
    n = pts.shape[0]
    if ts.shape[0] != n or w.shape[0] != n:
        raise ValueError(f"timestamps/weights must be ({n},), got {tuple(ts.shape)} / {tuple(w.shape)}")
    rg = None if ring is None else _dev(ring, torch.uint8).reshape(-1)
    tg = None if tag is None else _dev(tag, torch.uint8).reshape(-1)
    cap = int(n_points_cap)
    if cap < 1:
        raise ValueError("n_points_cap must be >= 1")
    po, to, wo = _empty(cap, 3), _empty(cap), _empty(cap)
    ro, go = _empty(cap, dtype=torch.uint8), _empty(cap, dtype=torch.uint8)
    sc, ws = _empty(8), _empty(4096)
    h = _h()
    _abi.check(_lib.gcl_point_budget_resample(h, _stream(), _ptr(pts), _ptr(ts), _ptr(w), _ptr(rg), _ptr(tg), n, cap,
                                              constants.GC_EPS_MASS, _ptr(po), _ptr(to), _ptr(wo), _ptr(ro), _ptr(go),
                                              _ptr(sc), _ptr(ws)), h)
    s = _host(sc)
    total_mass_in, ess = float(s[3]), float(s[6])
    result = PointBudgetResult(points=po, timestamps=to, weights=wo, ring=ro, tag=go, n_input=int(s[0]), n_output=int(s[1]),
                               total_mass_in=total_mass_in, total_mass_out=total_mass_in)
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["PointBudgetResample"],
        support=SupportCert(ess_total=ess, support_frac=float(min(1.0, cap / (n + constants.GC_EPS_MASS)))),
        influence=InfluenceCert.identity().with_overrides(
            mass_epsilon_ratio=constants.GC_EPS_MASS / (total_mass_in + constants.GC_EPS_MASS)))
    return result, cert, ExpectedEffect(objective_name="predicted_ess", predicted=ess, realized=None)


# ----------------------------------------------------------------------------------------------------------------
# DeskewConstantTwist
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class DeskewConstantTwistResult:
    """pkg/backend/operators/deskew_constant_twist.py:23-28."""
    points: torch.Tensor
    timestamps: torch.Tensor
    weights: torch.Tensor
    ess_imu: float


def deskew_constant_twist(points, timestamps, weights, scan_start_time: float, scan_end_time: float, xi_body, ess_imu: float,
                          chart_id: str, anchor_id: str) -> Tuple[DeskewConstantTwistResult, CertBundle, ExpectedEffect]:
    """pkg/backend/operators/deskew_constant_twist.py:72-117."""
    pts, ts, w = _dev(points).reshape(-1, 3), _dev(timestamps).reshape(-1), _dev(weights).reshape(-1)
    n = pts.shape[0]
    xi = np.asarray(xi_body.detach().cpu().numpy() if isinstance(xi_body, torch.Tensor) else xi_body, dtype=np.float64).reshape(-1)
    if xi.shape[0] != 6:
        raise ValueError(f"xi_body must be (6,), got {xi.shape}")
    po, wo, sc, ws = _empty(n, 3), _empty(n), _empty(4), _empty(4096)
    h = _h()
    _abi.check(_lib.gcl_deskew_constant_twist(h, _stream(), _ptr(pts), _ptr(ts), _ptr(w), n, float(scan_start_time),
                                              float(scan_end_time), _abi.dvec(xi), constants.GC_EPS_MASS, _ptr(po), _ptr(wo),
                                              _ptr(sc), _ptr(ws)), h)
    retained = float(_host(sc)[2])
    result = DeskewConstantTwistResult(points=po, timestamps=ts, weights=wo, ess_imu=float(ess_imu))
    cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id,
                                   support=SupportCert(ess_total=float(ess_imu), support_frac=retained),
                                   influence=InfluenceCert.identity())
    return result, cert, ExpectedEffect(objective_name="deskew_variance_reduction_proxy", predicted=0.0, realized=None)


def point_directions(points, lidar_origin, eps_mass: float = constants.GC_EPS_MASS) -> torch.Tensor:
    """Ray directions d = (p - o) / (|p - o| + eps), pkg/backend/pipeline.py:591-593."""
    pts = _dev(points).reshape(-1, 3)
    o = np.asarray(lidar_origin, dtype=np.float64).reshape(3)
    out = _empty(pts.shape[0], 3)
    h = _h()
    _abi.check(_lib.gcl_ray_directions(h, _stream(), _ptr(pts), pts.shape[0], _abi.dvec(o), eps_mass, _ptr(out)), h)
    return out


# ----------------------------------------------------------------------------------------------------------------
# BinSoftAssign / ScanBinMomentMatch / Kappa
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class BinSoftAssignResult:
    """archive/legacy_operators/binning.py:34-37."""
    responsibilities: torch.Tensor  # (N, B)


#: the materialised (N,B) operator refuses outputs above this many bytes (the fused pipeline never materialises it)
MAX_RESPONSIBILITY_BYTES = 8 << 30


def bin_soft_assign(point_directions, bin_directions, tau: float, chart_id: str = constants.GC_CHART_ID,
                    anchor_id: str = "initial") -> Tuple[BinSoftAssignResult, CertBundle, ExpectedEffect]:
    """archive/legacy_operators/binning.py:79-131.  `tau` is required: the reference's default constant is not in its
    tree."""
    d, b = _dev(point_directions).reshape(-1, 3), _dev(bin_directions).reshape(-1, 3)
    n, nb = d.shape[0], b.shape[0]
    if n * nb * 8 > MAX_RESPONSIBILITY_BYTES:
        raise ValueError(f"responsibilities ({n},{nb}) would need {n * nb * 8 / 2**30:.1f} GiB; use lidar_scan_evidence")
    resp, sc, ws = _empty(n, nb), _empty(4), _empty(4096)
    h = _h()
    _abi.check(_lib.gcl_bin_soft_assign(h, _stream(), _ptr(d), n, _ptr(b), nb, float(tau), constants.GC_EPS_MASS, _ptr(resp),
                                        _ptr(sc), _ptr(ws)), h)
    s = _host(sc)
    avg_entropy, max_resp = float(s[0]), float(s[1])
    cert = CertBundle.create_exact(chart_id=chart_id, anchor_id=anchor_id,
                                   support=SupportCert(ess_total=float(math.exp(avg_entropy)), support_frac=max_resp))
    return (BinSoftAssignResult(responsibilities=resp), cert,
            ExpectedEffect(objective_name="predicted_assignment_entropy", predicted=avg_entropy, realized=None))


@dataclass
class ScanBinStats:
    """archive/legacy_operators/binning.py:40-48."""
    N: torch.Tensor
    s_dir: torch.Tensor
    S_dir_scatter: torch.Tensor
    p_bar: torch.Tensor
    Sigma_p: torch.Tensor
    kappa_scan: torch.Tensor
    # raw position sums of the same pass (needed by the map update; not part of the reference dataclass)
    sum_p: Optional[torch.Tensor] = None
    sum_ppT: Optional[torch.Tensor] = None


def stats_views(rec: torch.Tensor, n_bins: int) -> ScanBinStats:
    """Typed views into one stats record (layout in include/gcl_b200.h)."""
    B = n_bins
    return ScanBinStats(N=rec[0:B], s_dir=rec[B:4 * B].view(B, 3), S_dir_scatter=rec[4 * B:13 * B].view(B, 3, 3),
                        p_bar=rec[13 * B:16 * B].view(B, 3), Sigma_p=rec[16 * B:25 * B].view(B, 3, 3),
                        kappa_scan=rec[25 * B:26 * B], sum_p=rec[26 * B:29 * B].view(B, 3),
                        sum_ppT=rec[29 * B:38 * B].view(B, 3, 3))


def _moment_match_cert(sc: np.ndarray, chart_id: str, anchor_id: str):
    ess = float(sc[_abi.SS["MM_ESS"]])
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["ScanBinMomentMatch"],
        support=SupportCert(ess_total=ess, support_frac=float(sc[_abi.SS["MM_SUPPORT_FRAC"]])),
        influence=InfluenceCert(lift_strength=0.0, psd_projection_delta=float(sc[_abi.SS["MM_PSD_DELTA"]]),
                                mass_epsilon_ratio=float(sc[_abi.SS["MM_MAX_EPS_RATIO"]]), anchor_drift_rho=0.0,
                                dt_scale=1.0, extrinsic_scale=1.0, trust_alpha=1.0))
    return cert, ExpectedEffect(objective_name="predicted_ess", predicted=ess, realized=None)


def scan_bin_moment_match(points, point_covariances, weights, responsibilities, point_lambda=None, direction_origin=None,
                          eps_psd: float = constants.GC_EPS_PSD, eps_mass: float = constants.GC_EPS_MASS,
                          chart_id: str = constants.GC_CHART_ID, anchor_id: str = "initial"
                          ) -> Tuple[ScanBinStats, CertBundle, ExpectedEffect]:
    """archive/legacy_operators/binning.py:212-324 (materialised responsibilities, optional per-point covariances)."""
    pts, w = _dev(points).reshape(-1, 3), _dev(weights).reshape(-1)
    n = pts.shape[0]
    r = _dev(responsibilities)
    if r.dim() != 2 or r.shape[0] != n:
        raise ValueError(f"responsibilities must be ({n}, B), got {tuple(r.shape)}")
    nb = r.shape[1]
    cov = None
    if point_covariances is not None:
        cov = _dev(point_covariances)
        if tuple(cov.shape) != (n, 3, 3):
            raise ValueError(f"point_covariances must be ({n},3,3), got {tuple(cov.shape)}")
    if direction_origin is None:
        o = np.zeros(3)
    else:
        o = np.asarray(direction_origin.detach().cpu().numpy() if isinstance(direction_origin, torch.Tensor)
                       else direction_origin, dtype=np.float64).reshape(-1)
        if o.shape[0] != 3:
            raise ValueError(f"direction_origin must be (3,), got {o.shape}")
    lam = None
    if point_lambda is not None:
        lam = _dev(point_lambda).reshape(-1)
        if lam.shape[0] != n:
            raise ValueError(f"point_lambda must be (N,), got {tuple(lam.shape)} for N={n}")
    if w.shape[0] != n:
        raise ValueError(f"weights must be ({n},), got {tuple(w.shape)}")
    h = _h()
    ws = _empty(int(_lib.gcl_moment_match_workspace_doubles(h, n, nb)))
    rec = _empty(int(_lib.gcl_stats_record_doubles(nb)))
    _abi.check(_lib.gcl_scan_bin_moment_match(h, _stream(), _ptr(pts), _ptr(cov), _ptr(w), _ptr(r), _ptr(lam), _abi.dvec(o), n,
                                              nb, float(eps_psd), float(eps_mass), _ptr(rec), _ptr(ws)), h)
    sc = _host(rec[38 * nb:])
    cert, effect = _moment_match_cert(sc, chart_id, anchor_id)
    return stats_views(rec, nb), cert, effect


# ----------------------------------------------------------------------------------------------------------------
# PointCloud2 ingest ("next" row)
# ----------------------------------------------------------------------------------------------------------------
def parse_pointcloud2_vlp16(msg=None, *, data=None, fields=None, point_step: Optional[int] = None, n_points: Optional[int] = None,
                            header_stamp_sec: float = 0.0, R_base_lidar=None, t_base_lidar=None, packed_t0: Optional[float] = None):
    """parse_pointcloud2_vlp16 (pkg/backend/backend_node.py:377-468) on the GPU: the raw message bytes are uploaded as they are
    (point_step bytes per point) and decoded there.  Pass a sensor_msgs/PointCloud2-like ``msg`` (``data``, ``fields`` with
    ``name`` / ``offset`` / ``datatype``, ``point_step``, ``width``, ``height``, ``header.stamp``) or the pieces by keyword
    (``fields`` as (name, offset, datatype) triples).  With ``R_base_lidar`` / ``t_base_lidar`` the points come back in the
    base frame (backend_node.py:1675-1681).  Returns ``(points, timestamps, weights, ring, tag)`` like the reference; with
    ``packed_t0`` also the 20 B/point layout ``(xyzt float4, w float)`` as a sixth and seventh element."""
    if msg is not None:
        data, point_step, n_points = msg.data, int(msg.point_step), int(msg.width) * int(msg.height)
        fields = [(f.name, int(f.offset), int(f.datatype)) for f in msg.fields]
        header_stamp_sec = msg.header.stamp.sec + msg.header.stamp.nanosec * 1e-9
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
    n = int(n_points)
    dev = _device()
    raw = torch.frombuffer(bytearray(bytes(data)[:n * point_step]), dtype=torch.uint8).to(dev) if n > 0 else torch.zeros(0, dtype=torch.uint8, device=dev)
    pts, ts, w = _empty(n, 3), _empty(n), _empty(n)
    ring, tag = _empty(n, dtype=torch.uint8), _empty(n, dtype=torch.uint8)
    xyzt = _empty(n, 4, dtype=torch.float32) if packed_t0 is not None else None
    wf = _empty(n, dtype=torch.float32) if packed_t0 is not None else None
    ws = _empty(2)
    R = None if R_base_lidar is None else np.ascontiguousarray(np.asarray(R_base_lidar, dtype=np.float64).reshape(9))
    t = None if t_base_lidar is None else np.ascontiguousarray(np.asarray(t_base_lidar, dtype=np.float64).reshape(3))
    h = _h()
    _abi.check(_lib.gcl_pointcloud2_ingest(h, _stream(), _ptr(raw), n, int(point_step), off.ctypes.data, dt.ctypes.data,
                                           float(header_stamp_sec), None if R is None else R.ctypes.data,
                                           None if t is None else t.ctypes.data, _ptr(pts), _ptr(ts), _ptr(w), _ptr(ring), _ptr(tag),
                                           _ptr(xyzt), _ptr(wf), 0.0 if packed_t0 is None else float(packed_t0), _ptr(ws)), h)
    return (pts, ts, w, ring, tag) if packed_t0 is None else (pts, ts, w, ring, tag, xyzt, wf)


# ----------------------------------------------------------------------------------------------------------------
# IMU preintegration -> deskew twist ("next" row)
# ----------------------------------------------------------------------------------------------------------------
IMU = dict(DELTA_POSE=0, DELTA_R=6, DELTA_P=15, DELTA_V=18, ESS=21, A_BODY_MEAN=22, A_WORLD_NOG_MEAN=25, A_WORLD_MEAN=28,
           DT_SUM=31, XI_BODY=32)
GCL_IMU_DOUBLES = 40


def _imu_launch(imu_stamps, imu_gyro, imu_accel, weights, window, rotvec_start_WB, gyro_bias, accel_bias, gravity_W,
                rotation_only: bool, want_weights: bool, jobs_dev: Optional[torch.Tensor] = None):
    t, gy, ac = _dev(imu_stamps).reshape(-1), _dev(imu_gyro).reshape(-1, 3), _dev(imu_accel).reshape(-1, 3)
    M = t.shape[0]
    if gy.shape[0] != M or ac.shape[0] != M:
        raise ValueError("imu_stamps, imu_gyro and imu_accel must have the same length")
    rv, bg, ba = _dev(rotvec_start_WB).reshape(-1, 3), _dev(gyro_bias).reshape(-1, 3), _dev(accel_bias).reshape(-1, 3)
    H = rv.shape[0]
    if bg.shape[0] != H or ba.shape[0] != H:
        raise ValueError("rotvec_start_WB, gyro_bias and accel_bias must have one row per hypothesis")
    w = None if weights is None else _dev(weights).reshape(-1)
    if w is not None and w.shape[0] != M:
        raise ValueError("weights must have one entry per IMU sample")
    t0, t1, sigma = window
    out = _empty(H, GCL_IMU_DOUBLES)
    w_out = _empty(M) if want_weights else None
    h = _h()
    _abi.check(_lib.gcl_imu_preintegrate(h, _stream(), _ptr(t), _ptr(gy), _ptr(ac), _ptr(w), M, float(t0), float(t1), float(sigma),
                                         _ptr(rv), _ptr(bg), _ptr(ba), _abi.dvec(np.asarray(gravity_W, dtype=np.float64).reshape(3)),
                                         1 if rotation_only else 0, H, _ptr(out), _ptr(w_out), _ptr(jobs_dev)), h)
    return out, w_out


def smooth_window_weights(imu_stamps, scan_start_time: float, scan_end_time: float, sigma: float) -> torch.Tensor:
    """pkg/backend/operators/imu_preintegration.py:19-43."""
    z = np.zeros((1, 3))
    M = int(np.asarray(imu_stamps.shape if hasattr(imu_stamps, "shape") else [len(imu_stamps)]).prod())
    _, w = _imu_launch(imu_stamps, np.zeros((M, 3)), np.zeros((M, 3)), None, (scan_start_time, scan_end_time, sigma), z, z, z,
                       np.zeros(3), False, True)
    return w


def preintegrate_imu_relative_pose(imu_stamps, imu_gyro, imu_accel, weights, rotvec_start_WB, gyro_bias, accel_bias, gravity_W):
    """preintegrate_imu_relative_pose_jax, pkg/backend/operators/imu_preintegration.py:46-147: the same nine outputs
    (delta_pose_se3, delta_R, delta_p, delta_v, ess, a_body_mean, a_world_nog_mean, a_world_mean, dt_eff_sum)."""
    out, _ = _imu_launch(imu_stamps, imu_gyro, imu_accel, weights, (0.0, 0.0, 1.0), rotvec_start_WB, gyro_bias, accel_bias,
                         gravity_W, False, False)
    o = out[0]
    return (o[0:6], o[6:15].view(3, 3), o[15:18], o[18:21], o[21], o[22:25], o[25:28], o[28:31], o[31])


preintegrate_imu_relative_pose_jax = preintegrate_imu_relative_pose   # the reference's name


def imu_deskew_twist(imu_stamps, imu_gyro, imu_accel, scan_start_time: float, scan_end_time: float, sigma: float,
                     rotvec_start_WB, gyro_bias, accel_bias, gravity_W, deskew_rotation_only: bool = False,
                     jobs_dev: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Window weights -> preintegration -> xi_body = se3_log(delta_pose) in one launch, batched over hypotheses (rows of
    rotvec_start_WB / gyro_bias / accel_bias); pkg/backend/pipeline.py:437-483.  With ``jobs_dev`` (a device array of H
    gcl_job records) the twists are written straight into the jobs of the fused LiDAR kernel."""
    out, w = _imu_launch(imu_stamps, imu_gyro, imu_accel, None, (scan_start_time, scan_end_time, sigma), rotvec_start_WB,
                         gyro_bias, accel_bias, gravity_W, deskew_rotation_only, True, jobs_dev)
    return {"record": out, "xi_body": out[:, IMU["XI_BODY"]:IMU["XI_BODY"] + 6], "delta_pose": out[:, 0:6], "ess": out[:, IMU["ESS"]],
            "dt_sum": out[:, IMU["DT_SUM"]], "weights": w}


def kappa_from_resultant_batch(R_bar, eps_r: float = constants.GC_EPS_R, d: int = 3, r0: float = constants.GC_KAPPA_BLEND_R0,
                               tau: float = constants.GC_KAPPA_BLEND_TAU) -> torch.Tensor:
    """pkg/backend/operators/kappa.py:130-169."""
    R = _dev(R_bar).reshape(-1)
    out = _empty(R.shape[0])
    h = _h()
    _abi.check(_lib.gcl_kappa_from_resultant(h, _stream(), _ptr(R), R.shape[0], float(eps_r), int(d), float(r0), float(tau),
                                             _ptr(out)), h)
    return out


# ----------------------------------------------------------------------------------------------------------------
# Bin atlas and map statistics
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class BinAtlas:
    """archive/bin_atlas.py:30-37."""
    dirs: torch.Tensor  # (B, 3)


def create_fibonacci_atlas(n_bins: int = constants.GC_B_BINS) -> BinAtlas:
    """archive/bin_atlas.py:40-75."""
    out = _empty(int(n_bins), 3)
    h = _h()
    _abi.check(_lib.gcl_fibonacci_atlas(h, _stream(), int(n_bins), constants.GC_EPS_MASS, _ptr(out)), h)
    return BinAtlas(dirs=out)


@dataclass
class MapBinStats:
    """archive/bin_atlas.py:83-101 (additive sufficient statistics, world frame)."""
    S_dir: torch.Tensor
    S_dir_scatter: torch.Tensor
    N_dir: torch.Tensor
    N_pos: torch.Tensor
    sum_p: torch.Tensor
    sum_ppT: torch.Tensor

    @classmethod
    def from_arrays(cls, d) -> "MapBinStats":
        return cls(**{k: _dev(d[k]) for k in ("S_dir", "S_dir_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT")})


def create_empty_map_stats(n_bins: int = constants.GC_B_BINS) -> MapBinStats:
    """archive/bin_atlas.py:116-134."""
    z = lambda *s: torch.zeros(s, dtype=torch.float64, device=_device())
    return MapBinStats(S_dir=z(n_bins, 3), S_dir_scatter=z(n_bins, 3, 3), N_dir=z(n_bins), N_pos=z(n_bins), sum_p=z(n_bins, 3),
                       sum_ppT=z(n_bins, 3, 3))


def update_map_stats(map_stats: MapBinStats, increments_S_dir, increments_S_dir_scatter, increments_N_dir, increments_N_pos,
                     increments_sum_p, increments_sum_ppT) -> MapBinStats:
    """archive/bin_atlas.py:137-163 (functional: returns new statistics)."""
    return MapBinStats(S_dir=map_stats.S_dir + _dev(increments_S_dir), S_dir_scatter=map_stats.S_dir_scatter + _dev(increments_S_dir_scatter),
                       N_dir=map_stats.N_dir + _dev(increments_N_dir), N_pos=map_stats.N_pos + _dev(increments_N_pos),
                       sum_p=map_stats.sum_p + _dev(increments_sum_p), sum_ppT=map_stats.sum_ppT + _dev(increments_sum_ppT))


def apply_forgetting(map_stats: MapBinStats, forgetting_factor: float = 0.99) -> MapBinStats:
    """archive/bin_atlas.py:232-257."""
    g = float(forgetting_factor)
    return MapBinStats(S_dir=g * map_stats.S_dir, S_dir_scatter=g * map_stats.S_dir_scatter, N_dir=g * map_stats.N_dir,
                       N_pos=g * map_stats.N_pos, sum_p=g * map_stats.sum_p, sum_ppT=g * map_stats.sum_ppT)


def _map_derived_buffer(map_stats: MapBinStats, eps_mass: float, eps_psd: float) -> torch.Tensor:
    nb = map_stats.N_dir.shape[0]
    out = _empty(16 * nb)
    h = _h()
    _abi.check(_lib.gcl_map_derived_stats(h, _stream(), _ptr(map_stats.S_dir), _ptr(map_stats.N_dir), _ptr(map_stats.N_pos),
                                          _ptr(map_stats.sum_p), _ptr(map_stats.sum_ppT), nb, float(eps_mass), float(eps_psd),
                                          _ptr(out)), h)
    return out


def compute_map_derived_stats(map_stats: MapBinStats, eps_mass: float = constants.GC_EPS_MASS, eps_psd: float = constants.GC_EPS_PSD):
    """archive/bin_atlas.py:203-224: returns (mu_dir, kappa, centroid, Sigma_c)."""
    nb = map_stats.N_dir.shape[0]
    out = _map_derived_buffer(map_stats, eps_mass, eps_psd)
    return (out[0:3 * nb].view(nb, 3), out[3 * nb:4 * nb], out[4 * nb:7 * nb].view(nb, 3), out[7 * nb:16 * nb].view(nb, 3, 3))


def map_update_inplace(map_stats: MapBinStats, stats_record: torch.Tensor, R_world, t_world, forgetting_factor: float = 0.99):
    """Forget, then add one scan's bin sums pushed to the world frame by (R, t) with t_z := 0 (rule declared in
    DESIGN.md; additive update archive/bin_atlas.py:137-163, forgetting :232-257).  Updates `map_stats` in place."""
    nb = map_stats.N_dir.shape[0]
    R, t = _dev(R_world).reshape(9), _dev(t_world).reshape(3)
    h = _h()
    _abi.check(_lib.gcl_map_update(h, _stream(), _ptr(stats_record), _ptr(R), _ptr(t), nb, float(forgetting_factor),
                                   _ptr(map_stats.S_dir), _ptr(map_stats.S_dir_scatter), _ptr(map_stats.N_dir),
                                   _ptr(map_stats.N_pos), _ptr(map_stats.sum_p), _ptr(map_stats.sum_ppT)), h)
    return map_stats


# ----------------------------------------------------------------------------------------------------------------
# Matrix Fisher rotation / planar translation evidence
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class ScatterMetrics:
    """archive/legacy_operators/matrix_fisher_evidence.py:54-63."""
    eigenvalues: torch.Tensor
    eigenvectors: torch.Tensor
    linearity: float
    planarity: float
    sphericity: float
    anisotropy: float
    effective_rank: float


@dataclass
class MatrixFisherResult:
    """archive/legacy_operators/matrix_fisher_evidence.py:66-75."""
    R_mf: torch.Tensor
    L_rot: torch.Tensor
    h_rot: torch.Tensor
    delta_rot: torch.Tensor
    svd_singular_values: torch.Tensor
    map_scatter_metrics: ScatterMetrics
    scan_scatter_metrics: ScatterMetrics


@dataclass
class PlanarTranslationResult:
    """archive/legacy_operators/matrix_fisher_evidence.py:402-410."""
    t_wls: torch.Tensor
    L_trans: torch.Tensor
    h_trans: torch.Tensor
    delta_trans: torch.Tensor
    xy_info_scale: float
    z_info_scale: float


def _scatter_metrics(ev: torch.Tensor, evh: np.ndarray, off_m: int, off_v: int) -> ScatterMetrics:
    m = evh[off_m:off_m + 8]
    return ScatterMetrics(eigenvalues=ev[off_m:off_m + 3], eigenvectors=ev[off_v:off_v + 9].view(3, 3), linearity=float(m[3]),
                          planarity=float(m[4]), sphericity=float(m[5]), anisotropy=float(m[6]), effective_rank=float(m[7]))


def mf_result_from_record(ev: torch.Tensor, evh: np.ndarray, chart_id: str, anchor_id: str):
    E = _abi.EV
    res = MatrixFisherResult(R_mf=ev[E["R_MF"]:E["R_MF"] + 9].view(3, 3), L_rot=ev[E["L_ROT"]:E["L_ROT"] + 9].view(3, 3),
                             h_rot=ev[E["H_ROT"]:E["H_ROT"] + 3], delta_rot=ev[E["DELTA_ROT"]:E["DELTA_ROT"] + 3],
                             svd_singular_values=ev[E["SVD_S"]:E["SVD_S"] + 3],
                             map_scatter_metrics=_scatter_metrics(ev, evh, E["MAP_METRICS"], E["MAP_EIGVECS"]),
                             scan_scatter_metrics=_scatter_metrics(ev, evh, E["SCAN_METRICS"], E["SCAN_EIGVECS"]))
    c = evh[E["ROT_CERT"]:E["ROT_CERT"] + 10]
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["MatrixFisherRotationEvidence"],
        conditioning=ConditioningCert(eig_min=float(c[0]), eig_max=float(c[1]), cond=float(c[2]), near_null_count=int(c[3])),
        mismatch=MismatchCert(nll_per_ess=float(c[4]), directional_score=float(c[5])),
        influence=InfluenceCert(lift_strength=0.0, psd_projection_delta=float(c[6]), mass_epsilon_ratio=float(c[7]),
                                anchor_drift_rho=0.0, dt_scale=1.0, extrinsic_scale=1.0, trust_alpha=1.0))
    return res, cert, ExpectedEffect(objective_name="predicted_rotation_nll", predicted=float(c[8]), realized=None)


def pt_result_from_record(ev: torch.Tensor, evh: np.ndarray, chart_id: str, anchor_id: str):
    E = _abi.EV
    c = evh[E["TRANS_CERT"]:E["TRANS_CERT"] + 12]
    res = PlanarTranslationResult(t_wls=ev[E["T_WLS"]:E["T_WLS"] + 3], L_trans=ev[E["L_TRANS"]:E["L_TRANS"] + 9].view(3, 3),
                                  h_trans=ev[E["H_TRANS"]:E["H_TRANS"] + 3], delta_trans=ev[E["DELTA_TRANS"]:E["DELTA_TRANS"] + 3],
                                  xy_info_scale=float(c[9]), z_info_scale=float(c[10]))
    cert = CertBundle.create_approx(
        chart_id=chart_id, anchor_id=anchor_id, triggers=["PlanarTranslationEvidence"],
        conditioning=ConditioningCert(eig_min=float(c[0]), eig_max=float(c[1]), cond=float(c[2]), near_null_count=int(c[3])),
        mismatch=MismatchCert(nll_per_ess=float(c[4]), directional_score=0.0),
        influence=InfluenceCert(lift_strength=0.0, psd_projection_delta=float(c[5]), mass_epsilon_ratio=float(c[6]),
                                anchor_drift_rho=0.0, dt_scale=1.0, extrinsic_scale=1.0, trust_alpha=1.0))
    return res, cert, ExpectedEffect(objective_name="predicted_translation_nll", predicted=float(c[7]), realized=None)


def _pack_scan_record(nb: int, **fields) -> torch.Tensor:
    rec = torch.zeros(int(_lib.gcl_stats_record_doubles(nb)), dtype=torch.float64, device=_device())
    v = stats_views(rec, nb)
    for k, t in fields.items():
        getattr(v, k).copy_(_dev(t).reshape(getattr(v, k).shape))
    return rec


def matrix_fisher_rotation_evidence(belief_pred, scan_s_dir, scan_S_dir_scatter, scan_N, map_S_dir, map_S_dir_scatter, map_N_dir,
                                    eps_psd: float = constants.GC_EPS_PSD, eps_lift: float = constants.GC_EPS_LIFT,
                                    eps_mass: float = constants.GC_EPS_MASS) -> Tuple[MatrixFisherResult, CertBundle, ExpectedEffect]:
    """archive/legacy_operators/matrix_fisher_evidence.py:264-394."""
    sN = _dev(scan_N).reshape(-1)
    nb = sN.shape[0]
    rec = _pack_scan_record(nb, N=sN, s_dir=scan_s_dir, S_dir_scatter=scan_S_dir_scatter)
    mS, mSc, mN = _dev(map_S_dir).reshape(nb, 3), _dev(map_S_dir_scatter).reshape(nb, 3, 3), _dev(map_N_dir).reshape(nb)
    pose = _dev(_pose6(belief_pred, eps_lift))
    ev = _empty(_abi.GCL_EVID_DOUBLES)
    h = _h()
    _abi.check(_lib.gcl_pose_evidence(h, _stream(), _ptr(rec), 0, _ptr(mS), _ptr(mSc), _ptr(mN), None, None, _ptr(pose), None, 1, nb,
                                      float(eps_psd), float(eps_mass), 1, 0, _ptr(ev)), h)
    return mf_result_from_record(ev, _host(ev), belief_pred.chart_id, belief_pred.anchor_id)


def planar_translation_evidence(belief_pred, scan_p_bar, scan_Sigma_p, scan_N, map_centroid, map_Sigma_c, map_N_pos,
                                map_S_dir_scatter, map_N_dir, R_hat, eps_psd: float = constants.GC_EPS_PSD,
                                eps_lift: float = constants.GC_EPS_LIFT, eps_mass: float = constants.GC_EPS_MASS
                                ) -> Tuple[PlanarTranslationResult, CertBundle, ExpectedEffect]:
    """archive/legacy_operators/matrix_fisher_evidence.py:502-671."""
    sN = _dev(scan_N).reshape(-1)
    nb = sN.shape[0]
    rec = _pack_scan_record(nb, N=sN, p_bar=scan_p_bar, Sigma_p=scan_Sigma_p)
    mder = torch.zeros(16 * nb, dtype=torch.float64, device=_device())
    mder[4 * nb:7 * nb].copy_(_dev(map_centroid).reshape(-1))
    mder[7 * nb:16 * nb].copy_(_dev(map_Sigma_c).reshape(-1))
    mSc, mNd, mNp = _dev(map_S_dir_scatter).reshape(nb, 3, 3), _dev(map_N_dir).reshape(nb), _dev(map_N_pos).reshape(nb)
    mS_dummy = torch.zeros(nb, 3, dtype=torch.float64, device=_device())
    Rh = _dev(R_hat).reshape(9)
    pose = _dev(_pose6(belief_pred, eps_lift))
    ev = _empty(_abi.GCL_EVID_DOUBLES)
    h = _h()
    _abi.check(_lib.gcl_pose_evidence(h, _stream(), _ptr(rec), 0, _ptr(mS_dummy), _ptr(mSc), _ptr(mNd), _ptr(mNp), _ptr(mder),
                                      _ptr(pose), _ptr(Rh), 1, nb, float(eps_psd), float(eps_mass), 0, 1, _ptr(ev)), h)
    return pt_result_from_record(ev, _host(ev), belief_pred.chart_id, belief_pred.anchor_id)


def build_combined_lidar_evidence_22d(mf_result: MatrixFisherResult, trans_result: PlanarTranslationResult):
    """archive/legacy_operators/matrix_fisher_evidence.py:729-756: translation block [0:3], rotation block [3:6]."""
    L = torch.zeros(constants.GC_D_Z, constants.GC_D_Z, dtype=torch.float64, device=mf_result.L_rot.device)
    hv = torch.zeros(constants.GC_D_Z, dtype=torch.float64, device=mf_result.L_rot.device)
    L[0:3, 0:3] = trans_result.L_trans
    hv[0:3] = trans_result.h_trans
    L[3:6, 3:6] = mf_result.L_rot
    hv[3:6] = mf_result.h_rot
    return L, hv
