"""BASELINE.json configs[1] and configs[2] at FULL size against the oracle (the per-scan recurrences -- map statistics with
forgetting, archive/bin_atlas.py:137-163,232-257, and the LiDAR measurement-noise IW block,
pkg/backend/operators/measurement_noise_iw_jax.py:60-100, called per scan at pkg/backend/backend_node.py:2106-2119 -- are where an
error could grow with the length of the stream), and the hard branches of the rotation evidence on the GPU:
so3_log near pi (pkg/common/geometry/se3_jax.py:330-366) and the det(U V^T) < 0 / H = 0 handling of the Matrix Fisher SVD
(archive/legacy_operators/matrix_fisher_evidence.py:215-222).  Tolerances: moments <= 1e-9, rotation <= 1e-8 rad.
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def g():
    import fl_slam_b200
    return fl_slam_b200


@pytest.fixture(scope="module")
def O():
    from oracle import gc_oracle
    return gc_oracle


def H(t):
    return t.detach().cpu().numpy()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0))


def chordal(Ra, Rb):
    """2 sin(theta/2) ~ theta: the arccos form bottoms out at 1.5e-8."""
    return float(np.linalg.norm(Ra - Rb) / math.sqrt(2.0))


# ----------------------------------------------------------------------------------------------------------------------
# configs[1]: 200 scans x ~30k points, cap 8192 (stride 4), 48 bins, map + IW state carried -- the whole stream
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("overlap", [False, True])
def test_configs1_stream_200_scans_against_oracle(g, O, overlap):
    import bench
    S, B, cap, tau = bench.N_SCANS, bench.N_BINS, bench.CAP, g.synth.DEFAULT_TAU
    scans, poses = bench.make_stream(S)
    bins = O.create_fibonacci_atlas(B)
    ref, ref_map, ref_iw = O.scan_stream(scans, poses, poses, n_points_cap=cap, bin_directions=bins, tau=tau,
                                         lidar_origin=scans[0].lidar_origin, iw_state=O.iw_lidar_init(1e-3), use_sigma_meas=True)
    pl = g.LidarEvidencePipeline(n_bins=B, tau=tau, precision="fp64", n_points_cap=cap, lidar_origin=scans[0].lidar_origin, bin_directions=bins)
    res = [pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans]
    ms, iw = g.create_empty_map_stats(B), pl.create_lidar_iw_state(1e-3)
    pd = torch.from_numpy(poses).cuda()
    xis = np.stack([s.xi_body for s in scans])
    ev = pl.run_stream(res, xis, pd, pd, ms, iw_state=iw, use_sigma_meas=True, overlap=overlap)
    pl.join_stream()
    torch.cuda.synchronize()
    st, evd = H(ev.stats), H(ev.evidence)
    worst = {"N": 0.0, "Sigma_p": 0.0, "R_mf": 0.0, "t_wls": 0.0, "L22": 0.0}
    for k in range(S):
        worst["N"] = max(worst["N"], rel_err(st[k, :B], ref[k]["scan_bins"]["N"]))
        worst["Sigma_p"] = max(worst["Sigma_p"], rel_err(st[k, 16 * B:25 * B].reshape(B, 3, 3), ref[k]["scan_bins"]["Sigma_p"]))
        if k > 0:   # scan 0 sees an empty map (H = 0): covered by test_matrix_fisher_zero_cross_covariance below
            worst["R_mf"] = max(worst["R_mf"], chordal(evd[k, 0:9].reshape(3, 3), ref[k]["matrix_fisher"]["R_mf"]))
            worst["t_wls"] = max(worst["t_wls"], rel_err(evd[k, 71:74], ref[k]["planar_translation"]["t_wls"]))
            worst["L22"] = max(worst["L22"], rel_err(evd[k, 123:123 + 484].reshape(22, 22), ref[k]["L_lidar"]))
    print("configs[1] full stream, worst over 200 scans:", worst)
    assert worst["N"] < 1e-9 and worst["Sigma_p"] < 1e-8, worst
    assert worst["R_mf"] < 1e-8, worst
    assert worst["t_wls"] < 1e-7 and worst["L22"] < 1e-7, worst
    for name in ("S_dir", "S_dir_scatter", "N_dir", "N_pos", "sum_p", "sum_ppT"):
        assert rel_err(H(getattr(ms, name)), ref_map[name]) < 1e-9, name
    iwh = H(iw)
    assert abs(iwh[0] / ref_iw["nu"] - 1) < 1e-12
    assert rel_err(iwh[1:].reshape(3, 3), ref_iw["Psi"]) < 1e-6


# ----------------------------------------------------------------------------------------------------------------------
# configs[2]: 64 hypotheses x 30k points (cap 32768, stride 1) x 48 bins against 64 oracle runs
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol_m,tol_r", [("fp64", 1e-9, 1e-8), ("fp32c", 1e-5, 1e-4)])
def test_configs2_64_hypotheses_against_oracle(g, O, precision, tol_m, tol_r):
    n, cap, B, Hn, tau = 30000, 32768, 48, 64, g.synth.DEFAULT_TAU
    sc = g.synth.make_scan(n, seed=1234 + 3)
    rng = np.random.default_rng(1234 + 3)
    xis = sc.xi_body[None] * (1 + 0.02 * rng.normal(size=(Hn, 6)))
    poses = 0.02 * rng.normal(size=(Hn, 6))
    bins = O.create_fibonacci_atlas(B)
    pb = O.point_budget_resample(sc.points, sc.timestamps, sc.weights, sc.ring, sc.tag, cap)
    dk0 = O.deskew_constant_twist(pb["points"], pb["timestamps"], pb["weights"], sc.scan_start_time, sc.scan_end_time, sc.xi_body)
    raw0 = O.fused_scan_raw_sums(dk0["points"], dk0["weights"], bins, tau, sc.lidar_origin)
    ms_h = g.synth.map_stats_from_scan_sums(raw0, *g.synth.known_map_transform(3), seed=3)
    ms = g.MapBinStats.from_arrays(ms_h)
    pl = g.LidarEvidencePipeline(n_bins=B, tau=tau, precision=precision, n_points_cap=cap, lidar_origin=sc.lidar_origin, bin_directions=bins)
    scan = pl.upload_scan(sc.points, sc.timestamps, sc.weights, sc.scan_start_time, sc.scan_end_time)
    ev = pl.run_resident(scan, xis, torch.from_numpy(poses).cuda(), ms)
    st, evd = H(ev.stats), H(ev.evidence)
    worst_m, worst_r, worst_t = 0.0, 0.0, 0.0
    md = O.compute_map_derived_stats(ms_h["S_dir"], ms_h["N_dir"], ms_h["N_pos"], ms_h["sum_p"], ms_h["sum_ppT"])
    for h in range(Hn):
        dk = O.deskew_constant_twist(pb["points"], pb["timestamps"], pb["weights"], sc.scan_start_time, sc.scan_end_time, xis[h])
        raw = O.fused_scan_raw_sums(dk["points"], dk["weights"], bins, tau, sc.lidar_origin)
        fin = O.scan_bin_finalize(raw)
        worst_m = max(worst_m, rel_err(st[h, :B], fin["N"]), rel_err(st[h, B:4 * B].reshape(B, 3), fin["s_dir"]),
                      rel_err(st[h, 26 * B:29 * B].reshape(B, 3), raw["sum_p"]), rel_err(st[h, 29 * B:38 * B].reshape(B, 3, 3), raw["sum_ppT"]))
        mf = O.matrix_fisher_rotation_evidence(poses[h], fin["s_dir"], fin["S_dir_scatter"], fin["N"], ms_h["S_dir"], ms_h["S_dir_scatter"], ms_h["N_dir"])
        worst_r = max(worst_r, chordal(evd[h, 0:9].reshape(3, 3), mf["R_mf"]))
        pt = O.planar_translation_evidence(poses[h], fin["p_bar"], fin["Sigma_p"], fin["N"], md["centroid"], md["Sigma_c"], ms_h["N_pos"],
                                           ms_h["S_dir_scatter"], ms_h["N_dir"], mf["R_mf"])
        worst_t = max(worst_t, rel_err(evd[h, 71:74], pt["t_wls"]))
    print(f"configs[2] 64 hypotheses {precision}: moments {worst_m:.2e}, R_mf {worst_r:.2e} rad, t_wls {worst_t:.2e}")
    assert worst_m < tol_m and worst_r < tol_r, (worst_m, worst_r)
    assert worst_t < (1e-7 if precision == "fp64" else 5e-3)


# ----------------------------------------------------------------------------------------------------------------------
# hard branches of the rotation evidence
# ----------------------------------------------------------------------------------------------------------------------
def _scan_and_map(g, O, F=None, seed=5):
    """Bin statistics of one synthetic scan and a map made of the same statistics moved by F (3x3; a rotation, or an improper
    matrix to force det(U V^T) < 0)."""
    sc = g.synth.make_scan(6000, seed=seed)
    bins = O.create_fibonacci_atlas(48)
    pb = O.point_budget_resample(sc.points, sc.timestamps, sc.weights, sc.ring, sc.tag, 8192)
    dk = O.deskew_constant_twist(pb["points"], pb["timestamps"], pb["weights"], sc.scan_start_time, sc.scan_end_time, sc.xi_body)
    raw = O.fused_scan_raw_sums(dk["points"], dk["weights"], bins, 0.1, sc.lidar_origin)
    fin = O.scan_bin_finalize(raw)
    F = np.eye(3) if F is None else F
    m = {"S_dir": fin["s_dir"] @ F.T, "S_dir_scatter": np.einsum("ij,bjk,lk->bil", F, fin["S_dir_scatter"], F), "N_dir": fin["N"].copy()}
    return fin, m


def _both(g, O, pose, fin, m):
    mf, cert, eff = g.matrix_fisher_rotation_evidence(g.PosePrediction(pose=pose, anchor_id="t"), fin["s_dir"], fin["S_dir_scatter"], fin["N"],
                                                      m["S_dir"], m["S_dir_scatter"], m["N_dir"])
    ref = O.matrix_fisher_rotation_evidence(pose, fin["s_dir"], fin["S_dir_scatter"], fin["N"], m["S_dir"], m["S_dir_scatter"], m["N_dir"])
    return mf, cert, ref


@pytest.mark.parametrize("offset", [0.0, 3e-8, -3e-8, 1e-3, -1e-3, 1e-5])
def test_so3_log_near_pi_on_the_gpu(g, O, offset):
    """delta = so3_log(R_pred^T R_mf) with the angle at pi - offset: |offset| < 1e-7 takes the softmax-mixture axis branch
    (se3_jax.py:346-363), larger offsets the general branch with sin(theta) ~ offset in the denominator."""
    Rstar = g.synth.rotation_from_rotvec(np.array([0.3, -0.2, 0.5]))
    fin, m = _scan_and_map(g, O, Rstar)
    axis = np.array([0.48, -0.6, 0.64])
    axis /= np.linalg.norm(axis)
    # R_pred^T R_mf = Rot(axis, pi - offset)  <=>  R_pred = R_mf Rot(axis, pi - offset)^T, with R_mf ~ Rstar
    R_pred = Rstar @ g.synth.rotation_from_rotvec(axis * (math.pi - offset)).T
    pose = np.concatenate([np.zeros(3), O.so3_log(R_pred)])
    mf, cert, ref = _both(g, O, pose, fin, m)
    th_ref = float(np.linalg.norm(ref["delta_rot"]))
    assert abs(th_ref - math.pi) < 2e-3
    d = H(mf.delta_rot)
    # the rotation vector itself is ill-conditioned at pi (sign of the axis, 1/sin): compare the ROTATIONS it stands for, and
    # the angle; away from the switch-over the vectors themselves
    assert chordal(g.synth.rotation_from_rotvec(d), g.synth.rotation_from_rotvec(ref["delta_rot"])) < 1e-6
    assert abs(float(np.linalg.norm(d)) - th_ref) < 1e-7
    if abs(offset) >= 1e-5:
        assert np.max(np.abs(d - ref["delta_rot"])) < 1e-12 / abs(offset) + 1e-11   # R_mf itself is good to ~cond(H) eps ~ 1e-14, amplified by pi / sin(theta)
    assert chordal(H(mf.R_mf), ref["R_mf"]) < 1e-8
    assert rel_err(H(mf.L_rot), ref["L_rot"]) < 1e-10


def test_matrix_fisher_reflection_branch(g, O):
    """Map directions = mirror image of the scan directions: the SVD of H gives det(U V^T) = -1 and the last column of U is
    flipped (matrix_fisher_evidence.py:215-222); L_rot still uses the unsigned singular values (:240-249)."""
    F = np.diag([1.0, 1.0, -1.0]) @ g.synth.rotation_from_rotvec(np.array([0.05, 0.02, -0.04]))
    fin, m = _scan_and_map(g, O, F)
    U, s, Vt = np.linalg.svd(O.matrix_fisher_rotation_evidence(np.zeros(6), fin["s_dir"], fin["S_dir_scatter"], fin["N"], m["S_dir"],
                                                               m["S_dir_scatter"], m["N_dir"])["H"])
    assert np.linalg.det(U @ Vt) < 0 and s[1] - s[2] > 1e-3 * s[0]       # the branch is taken and R_mf is well defined
    pose = np.array([0.0, 0.0, 0.0, 0.01, -0.02, 0.03])
    mf, cert, ref = _both(g, O, pose, fin, m)
    assert abs(np.linalg.det(H(mf.R_mf)) - 1.0) < 1e-12
    assert chordal(H(mf.R_mf), ref["R_mf"]) < 1e-8
    assert rel_err(H(mf.svd_singular_values), ref["svd_singular_values"]) < 1e-11
    assert rel_err(H(mf.L_rot), ref["L_rot"]) < 1e-10
    assert np.max(np.abs(H(mf.delta_rot) - ref["delta_rot"])) < 1e-9
    assert abs(cert.conditioning.eig_min - ref["cert_eig_min"]) < 1e-10 * ref["cert_eig_max"]


def test_matrix_fisher_zero_cross_covariance(g, O):
    """An empty map (scan 0 of every stream): H = 0 exactly.  LAPACK returns U = V = I for the zero matrix, so the reference
    yields R_mf = I, s = 0, L_raw = 0 -> L_rot = eps_psd I, delta = so3_log(R_pred^T)."""
    fin, m = _scan_and_map(g, O)
    m = {k: np.zeros_like(v) for k, v in m.items()}
    pose = np.array([0.0, 0.0, 0.0, 0.02, 0.01, -0.03])
    mf, cert, ref = _both(g, O, pose, fin, m)
    assert np.all(ref["H"] == 0.0) and chordal(ref["R_mf"], np.eye(3)) == 0.0
    assert np.all(np.isfinite(H(mf.R_mf))) and np.all(np.isfinite(H(mf.L_rot))) and np.all(np.isfinite(H(mf.delta_rot)))
    assert chordal(H(mf.R_mf), ref["R_mf"]) < 1e-12
    assert np.max(np.abs(H(mf.svd_singular_values))) < 1e-300
    assert rel_err(H(mf.L_rot), ref["L_rot"]) < 1e-12
    assert np.max(np.abs(H(mf.delta_rot) - ref["delta_rot"])) < 1e-12
