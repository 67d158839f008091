"""fl_slam_b200 -- B200-native (sm_100a) implementation of GC-SLAM v2's per-scan LiDAR evidence hot path, behind the
reference's operator contracts.  Importing the package loads libgcl_b200.so; there is no CPU fallback."""
from . import constants, synth  # noqa: F401
from .certificates import (CertBundle, ComputeCert, ConditioningCert, DeviceRuntimeCert, ExcitationCert,  # noqa: F401
                           ExpectedEffect, InfluenceCert, MismatchCert, SupportCert)
from . import _abi  # noqa: F401  (raises ImportError when the CUDA library is not built)
from .operators import (BinAtlas, BinSoftAssignResult, DeskewConstantTwistResult, MapBinStats, MatrixFisherResult,  # noqa: F401
                        PlanarTranslationResult, PointBudgetResult, PosePrediction, ScanBinStats, ScatterMetrics,
                        apply_forgetting, bin_soft_assign, build_combined_lidar_evidence_22d, compute_map_derived_stats,
                        create_empty_map_stats, create_fibonacci_atlas, deskew_constant_twist, imu_deskew_twist,
                        kappa_from_resultant_batch,
                        map_update_inplace, matrix_fisher_rotation_evidence, parse_pointcloud2_vlp16, planar_translation_evidence,
                        point_budget_resample,
                        point_directions, preintegrate_imu_relative_pose, preintegrate_imu_relative_pose_jax,
                        scan_bin_moment_match, smooth_window_weights, update_map_stats)
from .pipeline import LidarEvidencePipeline, LidarScanEvidence, ResidentScan, lidar_scan_evidence  # noqa: F401
from .manifest import runtime_manifest_entries  # noqa: F401

__version__ = "0.1.0"
