"""Numeric budgets and epsilons of the hot path (values of the reference's pkg/common/constants.py).

Line numbers refer to fl_ws/src/fl_slam_poc/fl_slam_poc/common/constants.py.  GC_B_BINS and GC_TAU_SOFT_ASSIGN are
NOT defined in the reference tree any more: B = 48 is recovered from CHANGELOG.md:492; tau has no
recoverable value, so GC_TAU_SOFT_ASSIGN is this build's declared default and every operator takes tau explicitly.
"""
GC_CHART_ID = "GC-RIGHT-01"      # :55
GC_K_HYP = 4                     # :62
GC_N_POINTS_CAP = 8192           # :64
GC_EPS_PSD = 1e-12               # :70
GC_EPS_LIFT = 1e-9               # :71
GC_EPS_MASS = 1e-12              # :72
GC_EPS_R = 1e-6                  # :73
GC_EPS_DEN = 1e-12               # :74
GC_KAPPA_BLEND_R0 = 0.8          # :95
GC_KAPPA_BLEND_TAU = 0.03        # :96
GC_TIME_WARP_SIGMA_FRAC = 0.1    # :143
GC_WEIGHT_FLOOR = 1e-12          # :256
GC_NONFINITE_SENTINEL = 1e6      # :257
GC_RANGE_WEIGHT_SIGMA = 0.25     # :260
GC_RANGE_WEIGHT_MIN_R = 0.5      # :261
GC_RANGE_WEIGHT_MAX_R = 50.0     # :262
GC_IW_NU_WEAK_ADD = 0.5          # :164
GC_LIDAR_SIGMA_MEAS = 0.01       # :210 (Kimera YAML overrides to 1e-3, config/gc_unified.yaml:16)
GC_IW_RHO_MEAS_LIDAR = 0.99      # :281
GC_IW_NU_MAX = 1000.0            # measurement_noise_iw_jax.py:66 default
GC_D_Z = 22
GC_B_BINS = 48                   # recovered, see module docstring
GC_TAU_SOFT_ASSIGN = 0.1         # DECLARED by this build, see module docstring
GC_POSE_EVIDENCE_BACKEND_BINS = "bins"
