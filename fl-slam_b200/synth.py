"""Seeded synthetic LiDAR scans of the shapes BASELINE.json names.

VLP-16-like sweep inside a box room, epoch-scale timestamps (so padded rows behave as in the
reference), the reference's range-window weights
(pkg/backend/backend_node.py:449-459: sigmoid((r-0.5)/0.25)*sigmoid((50-r)/0.25), floor 1e-12) and the
Kimera sensor origin (fl_ws/src/fl_slam_poc/config/gc_unified.yaml:18-21).  Pure NumPy; used by tests,
bench.py and smoke() to feed identical arrays to the CUDA path and to the CPU oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

LIDAR_ORIGIN_BASE = np.array([-0.065447, -0.100474, 0.108987], dtype=np.float64)
DEFAULT_TWIST = np.array([0.05, 0.01, 0.0, 0.002, -0.003, 0.08], dtype=np.float64)
T0_EPOCH = 1.7e9
SWEEP_SEC = 0.1
#: Declared soft-assign temperature.  The reference's GC_TAU_SOFT_ASSIGN is not in its tree
#:; 0.1 is about the nearest-neighbour similarity gap of the 48-bin atlas
#: (1 - cos 27 deg = 0.11).  Every report prints the value used.
DEFAULT_TAU = 0.1


@dataclass
class SyntheticScan:
    points: np.ndarray        # (n,3) f64, base frame
    timestamps: np.ndarray    # (n,) f64, epoch seconds
    weights: np.ndarray       # (n,) f64
    ring: np.ndarray          # (n,) u8
    tag: np.ndarray           # (n,) u8
    scan_start_time: float
    scan_end_time: float
    xi_body: np.ndarray       # (6,) [rho, phi] over the sweep
    lidar_origin: np.ndarray  # (3,)


def _sigmoid(x: np.ndarray) -> np.ndarray:
    return 1.0 / (1.0 + np.exp(-x))


def _room_range(origin: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """Distance along each ray to the walls of the box [-10,10]x[-6,6]x[-0.5,2.5]."""
    lo = np.array([-10.0, -6.0, -0.5])
    hi = np.array([10.0, 6.0, 2.5])
    with np.errstate(divide="ignore", invalid="ignore"):
        t_lo = (lo[None, :] - origin[None, :]) / dirs
        t_hi = (hi[None, :] - origin[None, :]) / dirs
    t_exit = np.where(dirs > 0, t_hi, np.where(dirs < 0, t_lo, np.inf))
    return np.min(t_exit, axis=1)


def make_scan(n_points: int, seed: int, *, dense: bool = False, t0: float = T0_EPOCH,
              xi_body: Optional[np.ndarray] = None, with_times: bool = True) -> SyntheticScan:
    """One synthetic sweep of ``n_points`` returns.  ``dense`` = jittered azimuth x elevation grid."""
    rng = np.random.default_rng(seed)
    n = int(n_points)
    if dense:
        n_el = max(16, int(round(np.sqrt(n / 8.0))))
        n_az = (n + n_el - 1) // n_el
        el_i, az_i = np.divmod(np.arange(n_el * n_az)[:n], n_az)
        elev = np.deg2rad(-15.0 + 30.0 * (el_i + rng.uniform(0, 1, n)) / n_el)
        azim = 2.0 * np.pi * (az_i + rng.uniform(0, 1, n)) / n_az
        ring = (el_i * 16 // n_el).astype(np.uint8)
        order = np.argsort(azim, kind="stable")            # sweep order = time order
        elev, azim, ring = elev[order], azim[order], ring[order]
    else:
        ring = rng.integers(0, 16, n).astype(np.uint8)
        elev = np.deg2rad(np.linspace(-15.0, 15.0, 16))[ring]
        azim = np.sort(rng.uniform(0.0, 2.0 * np.pi, n))
    dirs = np.stack([np.cos(elev) * np.cos(azim), np.cos(elev) * np.sin(azim), np.sin(elev)], axis=1)
    o = LIDAR_ORIGIN_BASE
    rng_m = np.clip(_room_range(o, dirs), 0.5, 50.0) + rng.normal(0.0, 0.03, n)
    pts = o[None, :] + rng_m[:, None] * dirs
    w = np.maximum(_sigmoid((rng_m - 0.5) / 0.25) * _sigmoid((50.0 - rng_m) / 0.25), 1e-12)
    if with_times:
        ts = t0 + azim / (2.0 * np.pi) * SWEEP_SEC
        t_end = t0 + SWEEP_SEC
    else:                                                    # scan without per-point times (backend_node.py:1782-1784)
        ts = np.full(n, t0)
        t_end = t0
    return SyntheticScan(points=np.ascontiguousarray(pts), timestamps=ts, weights=w, ring=ring,
                         tag=np.zeros(n, np.uint8), scan_start_time=float(t0), scan_end_time=float(t_end),
                         xi_body=np.array(DEFAULT_TWIST if xi_body is None else xi_body, dtype=np.float64),
                         lidar_origin=o.copy())


def rotation_from_rotvec(rv: np.ndarray) -> np.ndarray:
    rv = np.asarray(rv, dtype=np.float64)
    th = float(np.linalg.norm(rv))
    K = np.array([[0, -rv[2], rv[1]], [rv[2], 0, -rv[0]], [-rv[1], rv[0], 0]], dtype=np.float64)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def known_map_transform(seed: int = 0):
    """(R*, t*) of SURVEY.md 8d: |log R*| = 0.05 rad, t* = (0.2, -0.1, 0)."""
    rng = np.random.default_rng(9000 + seed)
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    return rotation_from_rotvec(0.05 * ax), np.array([0.2, -0.1, 0.0])


def map_stats_from_scan_sums(raw: Dict[str, np.ndarray], R: np.ndarray, t: np.ndarray, *, noise: float = 0.01,
                             seed: int = 0) -> Dict[str, np.ndarray]:
    """World-frame MapBinStats (archive/bin_atlas.py:83-101 fields) = the scan's raw bin sums moved by (R, t).

    ``raw`` holds N (B,), s_dir (B,3), S_dir_scatter (B,3,3), sum_p (B,3), sum_ppT (B,3,3) in the body frame.
    A per-bin multiplicative factor 1 + noise*N(0,1) perturbs the masses (applied consistently to all sums).
    """
    rng = np.random.default_rng(7000 + seed)
    N = np.asarray(raw["N"], dtype=np.float64)
    f = 1.0 + noise * rng.normal(size=N.shape[0])
    Rp = np.einsum("ij,bj->bi", R, raw["sum_p"])
    sum_p = Rp + N[:, None] * t[None, :]
    RppR = np.einsum("ij,bjk,lk->bil", R, raw["sum_ppT"], R)
    sum_ppT = (RppR + Rp[:, :, None] * t[None, None, :] + t[None, :, None] * Rp[:, None, :]
               + N[:, None, None] * (t[:, None] * t[None, :])[None])
    return {
        "S_dir": f[:, None] * np.einsum("ij,bj->bi", R, raw["s_dir"]),
        "S_dir_scatter": f[:, None, None] * np.einsum("ij,bjk,lk->bil", R, raw["S_dir_scatter"], R),
        "N_dir": f * N,
        "N_pos": f * N,
        "sum_p": f[:, None] * sum_p,
        "sum_ppT": f[:, None, None] * sum_ppT,
    }


def make_pointcloud2(scan: SyntheticScan, R_base_lidar=None, t_base_lidar=None, *, time_field: Optional[str] = "time_f32",
                     time_base: Optional[float] = None, header_stamp_sec: Optional[float] = None) -> Dict[str, object]:
    """The synthetic scan as the raw bytes of a VLP-16 PointCloud2 message (what pkg/backend/backend_node.py:377-468 parses):
    x, y, z float32 in the LiDAR frame, intensity float32, ring uint16, and a per-point time field --
    "time_f32" (the Kimera layout, 22 bytes per point: float32 seconds since ``time_base``), "t_f64" (float64 seconds since ``time_base``, 26 bytes; the reference parser reads any
    value above 1e6 as nanoseconds, so absolute epoch seconds cannot be carried), "t_u32ns" (uint32 nanoseconds since ``time_base``, 22 bytes; any value > 1e6 switches the reference
    parser to nanoseconds) or None (no time field, 18 bytes: every point takes the header stamp)."""
    n = scan.points.shape[0]
    R = np.eye(3) if R_base_lidar is None else np.asarray(R_base_lidar, dtype=np.float64).reshape(3, 3)
    t = np.zeros(3) if t_base_lidar is None else np.asarray(t_base_lidar, dtype=np.float64).reshape(3)
    p_lidar = (scan.points - t[None, :]) @ R                      # R^T (p_base - t)
    base = scan.scan_start_time if time_base is None else float(time_base)
    cols = [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("ring", "<u2")]
    if time_field == "time_f32":
        cols.append(("time", "<f4"))
    elif time_field == "t_f64":
        cols.append(("t", "<f8"))
    elif time_field == "t_u32ns":
        cols.append(("t", "<u4"))
    elif time_field is not None:
        raise ValueError(time_field)
    dt = np.dtype(cols)                                            # packed, unaligned offsets like the real message
    rec = np.zeros(n, dtype=dt)
    rec["x"], rec["y"], rec["z"] = p_lidar[:, 0], p_lidar[:, 1], p_lidar[:, 2]
    rec["intensity"] = 1.0
    rec["ring"] = scan.ring
    if time_field == "time_f32":
        rec["time"] = (scan.timestamps - base).astype(np.float32)
    elif time_field == "t_f64":
        rec["t"] = scan.timestamps - base
    elif time_field == "t_u32ns":
        rec["t"] = np.round((scan.timestamps - base) * 1e9).astype(np.uint32)
    code = {"<f4": 7, "<f8": 8, "<u2": 4, "<u4": 6}
    fields = [(name, int(dt.fields[name][1]), code[fmt]) for name, fmt in cols]
    stamp = (scan.scan_start_time - base if time_field is not None else scan.scan_start_time) \
        if header_stamp_sec is None else float(header_stamp_sec)
    return {"data": np.frombuffer(rec.tobytes(), dtype=np.uint8).copy(), "fields": fields, "point_step": int(dt.itemsize),
            "n_points": n, "header_stamp_sec": float(stamp)}


def make_imu_buffer(t0: float, t1: float, n_samples: int = 512, seed: int = 0) -> Dict[str, np.ndarray]:
    """A padded IMU buffer around the scan window [t0, t1] (pkg/backend/pipeline.py:437-483 takes M = 512 samples)."""
    rng = np.random.default_rng(seed)
    pad = 0.5 * (t1 - t0) + 0.02
    stamps = np.linspace(t0 - pad, t1 + pad, n_samples)
    gyro = np.array([0.02, -0.03, 0.8])[None, :] + 0.01 * rng.normal(size=(n_samples, 3))
    accel = np.array([0.3, -0.1, 9.81])[None, :] + 0.05 * rng.normal(size=(n_samples, 3))
    return {"stamps": stamps, "gyro": gyro, "accel": accel}
