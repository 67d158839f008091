"""ctypes binding of libgcl_b200.so (the C ABI in include/gcl_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every call that fails
raises with the library's own error text (reference policy: fail fast, docs/GC_SLAM.md:128-132).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# GCL_LIB: another build of the same library (A/B runs of kernel variants, trace builds); never a different implementation
LIB_PATH = os.environ.get("GCL_LIB") or os.path.join(_PKG_DIR, "libgcl_b200.so")

GCL_F64, GCL_F32C = 0, 1
GCL_LAYOUT_REF_F64, GCL_LAYOUT_PACKED_F32 = 0, 1
GCL_NM, GCL_NCERT, GCL_STATS_SCALARS, GCL_EVID_DOUBLES, GCL_MAX_BINS = 19, 8, 24, 640, 1024
GCL_MSG_INFO_HEADER = 8
GCL_MAX_RANKS, GCL_EXCHANGE_HANDLE_BYTES = 8, 64

# scalar tail of a stats record (enum GCL_SS_* in the header)
SS = dict(TOTAL_MASS_IN=0, TOTAL_MASS_SEL=1, MASS_SCALE=2, PB_ESS=3, N_SELECTED=4, STRIDE=5, DESKEW_RETAINED=6,
          AVG_ENTROPY=7, MAX_RESP=8, MM_ESS=9, MM_SUPPORT_FRAC=10, MM_PSD_DELTA=11, MM_MAX_EPS_RATIO=12,
          ENTROPY_VALID_ROWS=13, MAX_RESP_VALID_ROWS=14, SUM_W_OUT=15)
# offsets into an evidence record (enum GCL_EV_* in the header)
EV = dict(R_MF=0, L_ROT=9, H_ROT=18, DELTA_ROT=21, SVD_S=24, ROT_CERT=27, SCAN_METRICS=37, MAP_METRICS=45,
          SCAN_EIGVECS=53, MAP_EIGVECS=62, T_WLS=71, L_TRANS=74, H_TRANS=83, DELTA_TRANS=86, TRANS_CERT=89, H22=101,
          L22=123)


class GclJob(C.Structure):
    """Mirror of struct gcl_job."""
    _fields_ = [("points", C.c_void_p), ("timestamps", C.c_void_p), ("weights", C.c_void_p),
                ("point_lambda", C.c_void_p), ("n_raw", C.c_int64), ("n_points_cap", C.c_int32),
                ("sel_begin", C.c_int32), ("sel_end", C.c_int32), ("flags", C.c_int32),
                ("scan_start_time", C.c_double), ("scan_end_time", C.c_double), ("xi_body", C.c_double * 6)]


class GclError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
    pj = C.POINTER(GclJob)
    sig = {
        "gcl_abi_version": (i32, []),
        "gcl_create": (i32, [i32, C.POINTER(vp)]),
        "gcl_destroy": (None, [vp]),
        "gcl_last_error": (C.c_char_p, [vp]),
        "gcl_device_sm_count": (i32, [vp]),
        "gcl_launch_count": (i64, [vp]),
        "gcl_partial_record_doubles": (sz, [i32]),
        "gcl_stats_record_doubles": (sz, [i32]),
        "gcl_plan_ctas_per_job": (i32, [vp, i32, i64, i32, i32]),
        "gcl_scan_moments": (i32, [vp, vp, vp, pj, i32, i32, i32, i32, vp, i32, f64, C.POINTER(f64), f64, vp]),
        "gcl_reduce_finalize": (i32, [vp, vp, vp, i32, i64, i64, vp, pj, i32, i32, vp, f64, C.POINTER(f64), f64, f64,
                                      i32, vp, vp]),
        "gcl_map_derived_stats": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, f64, f64, vp]),
        "gcl_pose_evidence": (i32, [vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, i32, i32, f64, f64, i32, i32, vp]),
        "gcl_point_budget_resample": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, C.c_int32, f64, vp, vp, vp, vp, vp, vp, vp]),
        "gcl_deskew_constant_twist": (i32, [vp, vp, vp, vp, vp, i64, f64, f64, C.POINTER(f64), f64, vp, vp, vp, vp]),
        "gcl_ray_directions": (i32, [vp, vp, vp, i64, C.POINTER(f64), f64, vp]),
        "gcl_bin_soft_assign": (i32, [vp, vp, vp, i64, vp, i32, f64, f64, vp, vp, vp]),
        "gcl_moment_match_workspace_doubles": (sz, [vp, i64, i32]),
        "gcl_scan_bin_moment_match": (i32, [vp, vp, vp, vp, vp, vp, vp, C.POINTER(f64), i64, i32, f64, f64, vp, vp]),
        "gcl_kappa_from_resultant": (i32, [vp, vp, vp, i64, f64, i32, f64, f64, vp]),
        "gcl_fibonacci_atlas": (i32, [vp, vp, i32, f64, vp]),
        "gcl_map_update": (i32, [vp, vp, vp, vp, vp, i32, f64, vp, vp, vp, vp, vp, vp]),
        "gcl_scan_chain": (i32, [vp, vp, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, vp, i32, f64, f64, f64,
                                 f64, f64, vp]),
        "gcl_scan_chain_phase": (i32, [vp, vp, i32, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, vp, i32, f64, f64, f64,
                                       f64, f64, vp]),
        "gcl_scan_chain_workspace_doubles": (C.c_size_t, [i32, i32, i32]),
        "gcl_pointcloud2_ingest": (i32, [vp, vp, vp, i64, C.c_int32, vp, vp, f64, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64, vp]),
        "gcl_imu_preintegrate": (i32, [vp, vp, vp, vp, vp, vp, i32, f64, f64, f64, vp, vp, vp, vp, i32, i32, vp, vp, vp]),
        "gcl_pack_points_f32": (i32, [vp, vp, vp, vp, vp, i64, f64, vp, vp]),
        "gcl_exchange_create": (i32, [vp, i32, i32, sz, C.POINTER(vp), vp]),
        "gcl_exchange_connect": (i32, [vp, vp, vp]),
        "gcl_exchange_disconnect": (i32, [vp, vp]),
        "gcl_exchange_destroy": (None, [vp, vp]),
        "gcl_exchange_last_error": (C.c_char_p, [vp]),
        "gcl_exchange_error_flag": (i32, [vp]),
        "gcl_exchange_set_max_ctas": (i32, [vp, i32]),
        "gcl_exchange_set_lane": (i32, [vp, i32]),
        "gcl_scan_pipe_create": (i32, [vp, vp, i32, i32, i32, i32, C.POINTER(vp)]),
        "gcl_scan_pipe_destroy": (None, [vp]),
        "gcl_scan_pipe_last_error": (C.c_char_p, [vp]),
        "gcl_scan_pipe_submit": (i32, [vp, vp, i32, i32, vp, pj, i32, i32, i32, i32, vp, i32, f64, C.POINTER(f64), f64, f64, vp, vp, vp, vp,
                                       vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "gcl_scan_pipe_join": (i32, [vp, vp]),
        "gcl_scan_pipe_wait_set": (i32, [vp, i32, vp]),
        "gcl_scan_pipe_set_readback": (i32, [vp, i32, vp, vp]),
        "gcl_sharded_reduce_finalize": (i32, [vp, vp, vp, vp, i32, i64, i64, vp, pj, i32, i32, vp, f64, C.POINTER(f64), f64, f64, vp, vp]),
        "gcl_exchange_all_gather": (i32, [vp, vp, vp, vp, sz, vp]),
        "gcl_exchange_gathered": (vp, [vp, i32]),
        "gcl_scan_submit_host": (i32, [vp, i32, vp, vp, vp, i64, C.c_int32, f64, f64, vp, i32, vp, vp, i32, f64, i32, f64, f64]),
        "gcl_scan_submit_msg_host": (i32, [vp, i32, vp, i64, C.c_int32, vp, vp, f64, vp, vp, C.c_int32, vp, vp, vp, vp, i32, f64, vp, vp, vp,
                                           vp, i32, i32, vp, vp, i32, f64, i32, f64, f64]),
        "gcl_scan_msg_info_host": (i32, [vp, i32, vp, i64]),
        "gcl_host_register": (i32, [vp, vp, sz]),
        "gcl_host_unregister": (i32, [vp, vp]),
        "gcl_scan_evidence_begin_host": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, f64, vp, i32]),
        "gcl_scan_evidence_end_host": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "gcl_scan_evidence_host": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, f64, vp, i32, vp, vp]),
        "gcl_lidar_scan_evidence_host": (i32, [vp, vp, vp, vp, i64, C.c_int32, f64, f64, vp, vp, i32, vp, vp, i32, f64, i32,
                                               vp, vp, vp, vp, vp, vp, vp, f64, vp, i32, f64, f64, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = the .so does not export what the header declares
        fn.restype, fn.argtypes = res, args
    return lib, tuple(sig)


lib, EXPORTED = _load()
if lib.gcl_abi_version() != 1:
    raise ImportError("libgcl_b200.so ABI version mismatch")

_handles = {}
_lock = threading.Lock()


def handle(device: int = 0):
    """One library handle per (process, device, thread)."""
    key = (int(device), threading.get_ident())
    with _lock:
        h = _handles.get(key)
        if h is None:
            out = C.c_void_p()
            rc = lib.gcl_create(int(device), C.byref(out))
            if rc != 0:
                raise GclError(lib.gcl_last_error(None).decode())
            h = _handles[key] = out
        return h


def check(rc: int, h) -> None:
    if rc != 0:
        msg = lib.gcl_last_error(h).decode()
        if rc == -1:
            raise ValueError(msg)     # contract violation: the reference raises ValueError (binning.py:253-264)
        raise GclError(msg)


def dvec(values):
    arr = (C.c_double * len(values))(*[float(v) for v in values])
    return arr


def launch_count(device: int = 0) -> int:
    return int(lib.gcl_launch_count(handle(device)))
