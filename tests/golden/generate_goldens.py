#!/usr/bin/env python
"""Generate golden vectors by executing the REFERENCE'S OWN SOURCE (read from /root/reference, unmodified)
under the NumPy stand-in for jax in ``_jaxshim.py``.  Run in the build container only:

    python tests/golden/generate_goldens.py

Outputs (committed): tests/golden/ref_chain_*.npz, ref_pieces.npz, ref_certs.json.
Nothing here is imported by the product, by bench.py or by the -m gpu tests; /root/reference is not
needed at test time -- the tests read the committed .npz files.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

import _jaxshim  # noqa: E402


def _load_synth():
    spec = importlib.util.spec_from_file_location("gcl_synth", os.path.join(ROOT, "fl-slam_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["gcl_synth"] = m
    spec.loader.exec_module(m)
    return m


def _cert_numbers(cert):
    d = cert.to_dict()
    return {"exact": d["exact"], "approximation_triggers": d["approximation_triggers"],
            "frobenius_applied": d["frobenius_applied"], "conditioning": d["conditioning"], "support": d["support"],
            "mismatch": d["mismatch"], "influence": d["influence"],
            "total_trigger_magnitude": d["total_trigger_magnitude"]}


def run_chain(ref, synth, name, *, n_raw, cap, n_bins, tau, seed, with_times=True, pose_preds=None, use_lambda=False):
    """One full pass of the reference chain in pipeline order (README.md:105-120)."""
    jnp = sys.modules["jax.numpy"]
    sc = synth.make_scan(n_raw, seed, with_times=with_times)
    atlas = ref["bin_atlas"].create_fibonacci_atlas(n_bins)
    bins = np.asarray(atlas.dirs)
    out = {"in_points": sc.points, "in_timestamps": sc.timestamps, "in_weights": sc.weights, "in_ring": sc.ring,
           "in_tag": sc.tag, "in_scan_start_time": sc.scan_start_time, "in_scan_end_time": sc.scan_end_time,
           "in_xi_body": sc.xi_body, "in_lidar_origin": sc.lidar_origin, "in_n_points_cap": cap,
           "in_n_bins": n_bins, "in_tau": tau, "atlas_dirs": bins}
    certs = {}

    pb, c, e = ref["point_budget"].point_budget_resample(sc.points, sc.timestamps, sc.weights, sc.ring, sc.tag,
                                                         n_points_cap=cap)
    certs["point_budget"] = {"cert": _cert_numbers(c), "effect": e.to_dict(), "full_to_dict": c.to_dict()}
    for k in ("points", "timestamps", "weights", "ring", "tag"):
        out["pb_" + k] = np.asarray(getattr(pb, k))
    out["pb_scalars"] = np.array([pb.n_input, pb.n_output, pb.total_mass_in, pb.total_mass_out], dtype=np.float64)

    dk, c, e = ref["deskew_constant_twist"].deskew_constant_twist(
        pb.points, pb.timestamps, pb.weights, sc.scan_start_time, sc.scan_end_time, sc.xi_body, ess_imu=37.5,
        chart_id="GC-RIGHT-01", anchor_id="g")
    certs["deskew"] = {"cert": _cert_numbers(c), "effect": e.to_dict()}
    out["dk_points"], out["dk_weights"] = np.asarray(dk.points), np.asarray(dk.weights)

    # ray directions exactly as pkg/backend/pipeline.py:591-593
    rays = np.asarray(dk.points) - sc.lidar_origin[None, :]
    dirs = rays / (np.linalg.norm(rays, axis=1, keepdims=True) + ref["constants"].GC_EPS_MASS)
    out["dirs"] = dirs

    sa, c, e = ref["binning"].bin_soft_assign(dirs, bins, tau=tau)
    certs["soft_assign"] = {"cert": _cert_numbers(c), "effect": e.to_dict()}
    resp = np.asarray(sa.responsibilities)
    out["sa_row_sums"] = resp.sum(axis=1)
    out["sa_resp_head"] = resp[:64].copy()          # a slice is enough to pin the layout; full (N,B) is re-derived
    out["sa_resp_colsum"] = resp.sum(axis=0)

    lam = None
    if use_lambda:
        lam = 0.5 + np.random.default_rng(seed + 77).uniform(0, 1, cap)
        out["in_point_lambda"] = lam
    mm, c, e = ref["binning"].scan_bin_moment_match(
        points=dk.points, point_covariances=jnp.zeros((cap, 3, 3)), weights=dk.weights,
        responsibilities=sa.responsibilities, point_lambda=lam, direction_origin=sc.lidar_origin)
    certs["moment_match"] = {"cert": _cert_numbers(c), "effect": e.to_dict()}
    for k in ("N", "s_dir", "S_dir_scatter", "p_bar", "Sigma_p", "kappa_scan"):
        out["mm_" + k] = np.asarray(getattr(mm, k))

    # map = the scan's raw sums moved by the known (R*, t*); raw sums re-derived from the ref outputs
    w_r = (np.asarray(dk.weights) * (1.0 if lam is None else lam))[:, None] * resp
    P = np.asarray(dk.points)
    raw = {"N": w_r.sum(0), "s_dir": w_r.T @ dirs, "S_dir_scatter": np.einsum("nb,ni,nj->bij", w_r, dirs, dirs),
           "sum_p": w_r.T @ P, "sum_ppT": np.einsum("nb,ni,nj->bij", w_r, P, P)}
    Rstar, tstar = synth.known_map_transform(seed)
    ms = synth.map_stats_from_scan_sums(raw, Rstar, tstar, noise=0.01, seed=seed)
    out["map_Rstar"], out["map_tstar"] = Rstar, tstar
    for k, v in ms.items():
        out["map_" + k] = v
    BA = ref["bin_atlas"]
    mstats = BA.MapBinStats(S_dir=jnp.asarray(ms["S_dir"]), S_dir_scatter=jnp.asarray(ms["S_dir_scatter"]),
                            N_dir=jnp.asarray(ms["N_dir"]), N_pos=jnp.asarray(ms["N_pos"]),
                            sum_p=jnp.asarray(ms["sum_p"]), sum_ppT=jnp.asarray(ms["sum_ppT"]))
    mu_dir, kappa_map, centroid, Sigma_c = BA.compute_map_derived_stats(mstats)
    out["md_mu_dir"], out["md_kappa"] = np.asarray(mu_dir), np.asarray(kappa_map)
    out["md_centroid"], out["md_Sigma_c"] = np.asarray(centroid), np.asarray(Sigma_c)
    fg = BA.apply_forgetting(mstats, 0.99)
    out["forget_S_dir"], out["forget_sum_ppT"] = np.asarray(fg.S_dir), np.asarray(fg.sum_ppT)

    belief_mod = ref["belief"]
    MF = ref["matrix_fisher_evidence"]
    pose_preds = pose_preds if pose_preds is not None else [np.zeros(6)]
    out["in_pose_preds"] = np.asarray(pose_preds, dtype=np.float64)
    hyp = {k: [] for k in ("R_mf", "L_rot", "h_rot", "delta_rot", "svd_s", "t_wls", "L_trans", "h_trans",
                           "delta_trans", "L22", "h22", "scan_scatter_eigs", "map_scatter_eigs", "xyz_info")}
    certs["matrix_fisher"], certs["planar_translation"] = [], []
    for pose in pose_preds:
        belief = belief_mod.BeliefGaussianInfo.create_identity_prior(anchor_id="g", stamp_sec=0.0)
        belief.X_anchor = jnp.asarray(pose, dtype=jnp.float64)   # zero increment => mean_world_pose == X_anchor o Exp(0)
        out.setdefault("pose_world", [])
        out["pose_world"].append(np.asarray(belief.mean_world_pose()))
        mf, c, e = MF.matrix_fisher_rotation_evidence(belief, mm.s_dir, mm.S_dir_scatter, mm.N, mstats.S_dir,
                                                      mstats.S_dir_scatter, mstats.N_dir)
        certs["matrix_fisher"].append({"cert": _cert_numbers(c), "effect": e.to_dict(),
                                       "scan_metrics": {k: getattr(mf.scan_scatter_metrics, k) for k in
                                                        ("linearity", "planarity", "sphericity", "anisotropy", "effective_rank")},
                                       "map_metrics": {k: getattr(mf.map_scatter_metrics, k) for k in
                                                       ("linearity", "planarity", "sphericity", "anisotropy", "effective_rank")}})
        pt, c, e = MF.planar_translation_evidence(belief, mm.p_bar, mm.Sigma_p, mm.N, centroid, Sigma_c, mstats.N_pos,
                                                  mstats.S_dir_scatter, mstats.N_dir, mf.R_mf)
        certs["planar_translation"].append({"cert": _cert_numbers(c), "effect": e.to_dict()})
        L22, h22 = MF.build_combined_lidar_evidence_22d(mf, pt)
        for k, v in (("R_mf", mf.R_mf), ("L_rot", mf.L_rot), ("h_rot", mf.h_rot), ("delta_rot", mf.delta_rot),
                     ("svd_s", mf.svd_singular_values), ("t_wls", pt.t_wls), ("L_trans", pt.L_trans),
                     ("h_trans", pt.h_trans), ("delta_trans", pt.delta_trans), ("L22", L22), ("h22", h22),
                     ("scan_scatter_eigs", mf.scan_scatter_metrics.eigenvalues),
                     ("map_scatter_eigs", mf.map_scatter_metrics.eigenvalues),
                     ("xyz_info", np.array([pt.xy_info_scale, pt.z_info_scale]))):
            hyp[k].append(np.asarray(v))
    out["pose_world"] = np.stack(out["pose_world"])
    for k, v in hyp.items():
        out["hyp_" + k] = np.stack(v)
    np.savez_compressed(os.path.join(HERE, f"ref_chain_{name}.npz"), **out)
    return certs


def run_pieces(ref):
    """Known-answer vectors for the small pieces (kappa, SO3/SE3, PSD projection, atlas, inv-mass)."""
    rng = np.random.default_rng(2024)
    out = {}
    Rv = np.concatenate([np.linspace(0.0, 1.0, 41), [0.1, 0.3, 0.5, 0.7, 0.85, 0.99, 0.999999, 1.5, -0.2]])
    out["kappa_in"] = Rv
    out["kappa_out"] = np.asarray(ref["kappa"].kappa_from_resultant_batch(Rv))
    out["kappa_scalar_out"] = np.array([ref["kappa"].kappa_from_resultant_v2(float(r))[0].kappa for r in Rv])
    for B in (16, 48, 256, 1024):
        out[f"atlas_{B}"] = np.asarray(ref["bin_atlas"].create_fibonacci_atlas(B).dirs)
    se3 = ref["se3_jax"]
    omegas = np.concatenate([rng.normal(size=(12, 3)) * np.array([[1e-9], [1e-4], [0.01], [0.1], [0.5], [1.0], [1.5], [2.5],
                                                                 [3.0], [0.3], [0.7], [2.0]]),
                             [[np.pi - 0.01, 0, 0], [0, 0, 0], [1e-8, 0, 0], [0, np.pi - 1e-9, 0]]])
    out["so3_omega"] = omegas
    out["so3_exp"] = np.stack([np.asarray(se3.so3_exp(o)) for o in omegas])
    out["so3_log_of_exp"] = np.stack([np.asarray(se3.so3_log(se3.so3_exp(o))) for o in omegas])
    xis = rng.normal(size=(10, 6)) * np.array([[1e-9], [1e-3], [0.05], [0.2], [0.5], [1.0], [2.0], [0.1], [0.01], [3.0]])
    out["se3_xi"] = xis
    out["se3_exp"] = np.stack([np.asarray(se3.se3_exp(x)) for x in xis])
    poses_a, poses_b = rng.normal(size=(6, 6)) * 0.7, rng.normal(size=(6, 6)) * 0.7
    out["se3_a"], out["se3_b"] = poses_a, poses_b
    out["se3_compose"] = np.stack([np.asarray(se3.se3_compose(a, b)) for a, b in zip(poses_a, poses_b)])
    mats = []
    for i in range(8):
        A = rng.normal(size=(3, 3))
        mats.append(A @ A.T - (0.3 if i % 2 else 0.0) * np.eye(3) + (1e-3 * rng.normal(size=(3, 3)) if i % 3 == 0 else 0))
    mats.append(np.diag([1.0, -1000.0, 1.0]))
    mats.append(np.zeros((3, 3)))
    mats.append(1e-14 * np.eye(3))
    mats = np.stack(mats)
    out["psd_in"] = mats
    res = [ref["primitives"].domain_projection_psd_core(M, 1e-12) for M in mats]
    out["psd_out"] = np.stack([np.asarray(r[0]) for r in res])
    out["psd_cert"] = np.stack([np.asarray(r[1]) for r in res])
    m = np.array([0.0, 1e-13, 1e-6, 1.0, 123.4, -1e-13])
    inv, ratio = ref["primitives"].inv_mass_core(m, 1e-12)
    out["invmass_in"], out["invmass_inv"], out["invmass_ratio"] = m, np.asarray(inv), np.asarray(ratio)
    t = np.linspace(-0.05, 0.15, 33) + 1.7e9
    out["sww_t"] = t
    out["sww_w"] = np.asarray(ref["imu_preintegration"].smooth_window_weights(t, 1.7e9, 1.7e9 + 0.1, 0.01))
    # LiDAR measurement-noise inverse-Wishart block: three successive updates from random per-bin residuals
    jnp = sys.modules["jax.numpy"]
    st = ref["mn_iw_struct"].create_datasheet_measurement_noise_state(lidar_sigma_meas=1e-3)
    out["iw_nu0"], out["iw_Psi0"] = np.asarray(st.nu), np.asarray(st.Psi_blocks)
    res_all, w_all, nus, Psis, Sigs = [], [], [], [], []
    for k in range(3):
        res = rng.normal(size=(48, 3)) * np.array([0.02, 0.03, 0.005]) * (1 + k)
        w = rng.uniform(0.0, 2.0, 48)
        dPsi, dnu = ref["mn_iw"].lidar_meas_iw_suffstats_from_translation_residuals_jax(res, w)
        st, _ = ref["mn_iw"].measurement_noise_apply_suffstats_jax(st, dPsi, dnu)
        res_all.append(res); w_all.append(w); nus.append(np.asarray(st.nu)); Psis.append(np.asarray(st.Psi_blocks))
        Sigs.append(np.asarray(ref["mn_iw"].measurement_noise_mean_jax(st, 2)))
    out["iw_res"], out["iw_w"] = np.stack(res_all), np.stack(w_all)
    out["iw_nu"], out["iw_Psi"], out["iw_Sigma_lidar"] = np.stack(nus), np.stack(Psis), np.stack(Sigs)
    np.savez_compressed(os.path.join(HERE, "ref_pieces.npz"), **out)


def run_imu(ref):
    """IMU preintegration -> deskew twist: the reference's own smooth_window_weights,
    preintegrate_imu_relative_pose_jax (imu_preintegration.py:46-147) and se3_log (se3_jax.py:221-256) on seeded IMU streams."""
    rng = np.random.default_rng(77)
    pre, se3 = ref["imu_preintegration"], ref["se3_jax"]
    out = {}
    cases = {"a": dict(M=200, rate=200.0, w_scale=0.5, lead=0.02), "b": dict(M=512, rate=400.0, w_scale=1.5, lead=0.3),
             "c": dict(M=37, rate=100.0, w_scale=0.05, lead=0.0)}
    for name, c in cases.items():
        M = c["M"]
        t0 = 1.7e9
        stamps = t0 - c["lead"] + np.arange(M) / c["rate"] + rng.uniform(0, 2e-4, M)
        gyro = c["w_scale"] * rng.normal(size=(M, 3)) + np.array([0.0, 0.0, 0.8])
        accel = rng.normal(size=(M, 3)) * 0.5 + np.array([0.0, 0.0, 9.81])
        rotvec0 = 0.3 * rng.normal(size=3)
        bg, ba = 0.01 * rng.normal(size=3), 0.05 * rng.normal(size=3)
        grav = np.array([0.0, 0.0, -9.81])
        w = np.asarray(pre.smooth_window_weights(stamps, t0, t0 + 0.1, 0.01))
        res = pre.preintegrate_imu_relative_pose_jax(stamps, gyro, accel, w, rotvec0, bg, ba, grav)
        xi = np.asarray(se3.se3_log(res[0]))
        out.update({f"{name}_stamps": stamps, f"{name}_gyro": gyro, f"{name}_accel": accel, f"{name}_rotvec0": rotvec0,
                    f"{name}_bg": bg, f"{name}_ba": ba, f"{name}_grav": grav, f"{name}_t0": t0, f"{name}_t1": t0 + 0.1,
                    f"{name}_sigma": 0.01, f"{name}_w": w, f"{name}_xi": xi})
        for k, nm in enumerate(("delta_pose", "delta_R", "delta_p", "delta_v", "ess", "a_body_mean", "a_world_nog_mean",
                                "a_world_mean", "dt_sum")):
            out[f"{name}_{nm}"] = np.asarray(res[k])
    # se3_log known answers, including the small-angle branch
    poses = rng.normal(size=(8, 6)) * np.array([[1e-9], [1e-5], [1e-3], [0.05], [0.3], [1.0], [2.5], [3.1]])
    out["se3log_in"] = poses
    out["se3log_out"] = np.stack([np.asarray(se3.se3_log(p)) for p in poses])
    np.savez_compressed(os.path.join(HERE, "ref_imu.npz"), **out)


def _load_pointcloud2_parser():
    """parse_pointcloud2_vlp16 and _pointfield_to_dtype cut out of the reference's backend_node.py (the module itself imports
    rclpy / sensor_msgs, which are not installed): the two function bodies are executed unmodified against a stand-in
    PointField (sensor_msgs/msg/PointField datatype codes) and the reference's own constants module."""
    import ast, importlib.util, types
    path = "/root/reference/fl_ws/src/fl_slam_poc/fl_slam_poc/backend/backend_node.py"
    src = open(path).read()
    tree = ast.parse(src)
    wanted = {"_pointfield_to_dtype", "parse_pointcloud2_vlp16"}
    chunks = [ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert len(chunks) == 2
    spec = importlib.util.spec_from_file_location("ref_constants", "/root/reference/fl_ws/src/fl_slam_poc/fl_slam_poc/common/constants.py")
    constants = importlib.util.module_from_spec(spec); spec.loader.exec_module(constants)
    PointField = types.SimpleNamespace(INT8=1, UINT8=2, INT16=3, UINT16=4, INT32=5, UINT32=6, FLOAT32=7, FLOAT64=8)
    from typing import Tuple
    ns = {"np": np, "constants": constants, "PointField": PointField, "PointCloud2": object, "Tuple": Tuple}
    exec("\n\n".join(chunks), ns)
    return ns["parse_pointcloud2_vlp16"]


def _make_msg(data, fields, point_step, n, stamp):
    import types
    return types.SimpleNamespace(width=n, height=1, point_step=point_step, data=bytes(data),
                                 fields=[types.SimpleNamespace(name=f[0], offset=f[1], datatype=f[2]) for f in fields],
                                 header=types.SimpleNamespace(stamp=types.SimpleNamespace(sec=int(stamp[0]), nanosec=int(stamp[1]))))


def run_pointcloud2():
    """PointCloud2 ingest: the reference's parse_pointcloud2_vlp16 (backend_node.py:377-468) and the base-frame
    transform (:1675-1681) on synthetic VLP-16 messages: the Kimera layout (x y z f32, intensity f32, ring u16, time f32,
    point_step 22), per-point time in nanoseconds as u32, a layout without per-point time, doubles, and non-finite points."""
    parse = _load_pointcloud2_parser()
    rng = np.random.default_rng(31)
    out = {}
    R_bl = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]) @ np.array([[1, 0, 0], [0, 0.9998, -0.02], [0, 0.02, 0.9998]])
    t_bl = np.array([-0.065447, -0.100474, 0.108987])
    layouts = {
        "kimera": dict(step=22, fields=[("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 12, 7), ("ring", 16, 4), ("time", 18, 7)]),
        "ns_u32": dict(step=24, fields=[("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("ring", 12, 2), ("t", 16, 6)]),
        "no_time": dict(step=16, fields=[("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("ring", 12, 4)]),
        "f64": dict(step=40, fields=[("x", 0, 8), ("y", 8, 8), ("z", 16, 8), ("ring", 24, 5), ("time", 32, 8)]),
    }
    npd = {1: "i1", 2: "u1", 3: "<i2", 4: "<u2", 5: "<i4", 6: "<u4", 7: "<f4", 8: "<f8"}
    for name, L in layouts.items():
        n = 777
        dt = np.dtype({"names": [f[0] for f in L["fields"]], "formats": [npd[f[2]] for f in L["fields"]],
                       "offsets": [f[1] for f in L["fields"]], "itemsize": L["step"]})
        arr = np.zeros(n, dtype=dt)
        rngm = rng.uniform(0.2, 60.0, n)
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        xyz = d * rngm[:, None]
        xyz[5] = [np.nan, 1.0, 2.0]; xyz[17] = [np.inf, -np.inf, 0.0]; xyz[40] = [0.0, 0.0, 0.0]
        arr["x"], arr["y"], arr["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        arr["ring"] = rng.integers(0, 16, n)
        if "intensity" in dt.names:
            arr["intensity"] = rng.uniform(0, 255, n)
        tf = [f for f in L["fields"] if f[0] in ("t", "time")]
        if tf:
            arr[tf[0][0]] = (rng.uniform(0, 0.1, n) * (1e9 if tf[0][2] == 6 else 1.0))
        raw = arr.tobytes()
        stamp = (1700000000, 123456789)
        pts, ts, w, ring, tag = parse(_make_msg(raw, L["fields"], L["step"], n, stamp))
        pts_base = (R_bl @ pts.T).T + t_bl[None, :]
        out.update({f"{name}_raw": np.frombuffer(raw, dtype=np.uint8), f"{name}_step": L["step"], f"{name}_n": n,
                    f"{name}_fields": np.array([[{"x": 0, "y": 1, "z": 2, "ring": 3, "t": 4, "time": 4, "intensity": 5}[f[0]], f[1], f[2]] for f in L["fields"]]),
                    f"{name}_stamp": np.array(stamp), f"{name}_points_lidar": pts, f"{name}_points_base": pts_base, f"{name}_t": ts,
                    f"{name}_w": w, f"{name}_ring": ring, f"{name}_tag": tag})
    out["R_base_lidar"], out["t_base_lidar"] = R_bl, t_bl
    np.savez_compressed(os.path.join(HERE, "ref_pointcloud2.npz"), **out)


def run_time_bounds():
    """Scan time bounds (pkg/backend/backend_node.py:1776-1786): the statements are inline in a ROS node method, so they are cut out
    of the reference source as TEXT (from the `time_bounds = jnp.array(` line to the line before `dt_raw = (`), dedented and
    executed unmodified on `timestamps` / `stamp_sec` with NumPy standing in for jnp."""
    import textwrap
    path = "/root/reference/fl_ws/src/fl_slam_poc/fl_slam_poc/backend/backend_node.py"
    lines = open(path).read().split("\n")
    a = next(i for i, l in enumerate(lines) if l.strip().startswith("time_bounds = jnp.array("))
    b = next(i for i in range(a, len(lines)) if lines[i].strip().startswith("dt_raw = ("))
    code = textwrap.dedent("\n".join(lines[a:b]))
    assert "scan_start_time" in code and "scan_end_time" in code and "1e-9" in code
    rng = np.random.default_rng(77)
    cases = []
    for k in range(12):
        base = [0.0, 12.25, 1.7e9][k % 3]
        if k % 4 == 3:
            ts = np.full(50, base + 0.05)                              # no per-point spread
        elif k == 5:
            ts = base + np.array([0.0, 4e-10, 9e-10])                  # spread below the 1e-9 threshold
        else:
            ts = base + np.sort(rng.uniform(0.0, 0.1, 200))
        for stamp in (base - 0.3, base + 0.04, base + 0.7):
            cases.append((ts, float(stamp)))
    out = {"n_cases": len(cases)}
    for i, (ts, stamp) in enumerate(cases):
        ns = {"jnp": np, "timestamps": ts, "stamp_sec": stamp, "record_device_to_host": lambda *a_, **k_: None}
        exec(code, ns)
        out[f"ts_{i}"] = ts
        out[f"io_{i}"] = np.array([stamp, ns["scan_start_time"], ns["scan_end_time"]])
    np.savez_compressed(os.path.join(HERE, "ref_time_bounds.npz"), **out)


def main():
    ref = _jaxshim.load_reference("/root/reference", n_bins=48, tau=0.1)
    synth = _load_synth()
    rng = np.random.default_rng(5)
    poses3 = [np.zeros(6), np.concatenate([[0.15, -0.05, 0.01], 0.04 * rng.normal(size=3)]),
              np.concatenate([[-0.3, 0.2, 0.0], 0.2 * rng.normal(size=3)])]
    certs = {}
    certs["stride3_pad"] = run_chain(ref, synth, "stride3_pad", n_raw=1500, cap=512, n_bins=48, tau=0.1, seed=1235,
                                     pose_preds=poses3)
    certs["exact_fit"] = run_chain(ref, synth, "exact_fit", n_raw=1024, cap=1024, n_bins=48, tau=0.05, seed=1236,
                                   use_lambda=True)
    certs["short_pad_b16"] = run_chain(ref, synth, "short_pad_b16", n_raw=300, cap=512, n_bins=16, tau=0.2, seed=1237)
    certs["no_times"] = run_chain(ref, synth, "no_times", n_raw=700, cap=256, n_bins=48, tau=0.1, seed=1238,
                                  with_times=False)
    certs["b256"] = run_chain(ref, synth, "b256", n_raw=600, cap=640, n_bins=256, tau=0.03, seed=1239)
    with open(os.path.join(HERE, "ref_certs.json"), "w") as f:
        json.dump(certs, f, indent=1, sort_keys=True, default=float)
    run_pieces(ref)
    run_imu(ref)
    run_pointcloud2()
    run_time_bounds()
    print("wrote goldens to", HERE)


if __name__ == "__main__":
    main()
