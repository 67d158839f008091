"""Runtime-manifest entries for the bin-based LiDAR evidence path.

The reference publishes a RuntimeManifest at start-up (pkg/backend/pipeline.py:1629-1793; required by
docs/GC_SLAM.md:849-859): a ``backends`` dict naming the implementation behind every operator slot plus the numeric
budgets.  A maintainer merges these entries into RuntimeManifest.backends / to_dict().
"""
from __future__ import annotations

from typing import Any, Dict

from . import constants


def runtime_manifest_entries(pipeline, n_gpus: int = 1, shard_mode: str = "none") -> Dict[str, Any]:
    impl = f"fl_slam_b200 (libgcl_b200.so, sm_100a CUDA; {pipeline.precision})"
    return {
        "pose_evidence_backend": constants.GC_POSE_EVIDENCE_BACKEND_BINS,
        "backends": {
            "point_budget": f"{impl} fused gather in gcl_scan_moments",
            "deskew": f"{impl} constant twist fused in gcl_scan_moments",
            "bin_soft_assign": f"{impl} fused softmax, responsibilities never materialised",
            "scan_bin_moment_match": f"{impl} gcl_scan_moments + gcl_reduce_finalize (fixed-order reduction)",
            "kappa": f"{impl} gcl_reduce_finalize",
            "domain_projection_psd": f"{impl} 3x3 cyclic Jacobi (gcl_math.cuh)",
            "lidar_evidence": f"{impl} gcl_pose_evidence (Matrix Fisher rotation + planar translation)",
            "map_update": f"{impl} gcl_map_update (bin statistics; forgetting)",
            "se3": f"{impl} gcl_math.cuh (so3_exp / so3_log / constant-twist deskew)",
        },
        "N_POINTS_CAP": pipeline.n_points_cap,
        "B_BINS": pipeline.n_bins,
        "tau_soft_assign": pipeline.tau,
        "precision_mode": pipeline.precision,
        "eps_psd": pipeline.eps_psd,
        "eps_mass": pipeline.eps_mass,
        "eps_lift": constants.GC_EPS_LIFT,
        "eps_r": constants.GC_EPS_R,
        "n_gpus": int(n_gpus),
        "shard_mode": shard_mode,
    }
