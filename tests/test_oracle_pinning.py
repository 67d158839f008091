"""Pin oracle/gc_oracle.py against the reference.

(1) tests/golden/ref_*.npz hold outputs of the reference's own source files executed under the NumPy
    stand-in for jax (tests/golden/generate_goldens.py); the oracle must reproduce them.
(2) The numeric tests the reference itself keeps for pieces of this path, re-expressed against the oracle:
    test/test_audit_invariants.py:103-119 (kappa smooth), :224-330 (SO3/SE3 round trips),
    :398-413 (kappa batch == scalar, rtol 1e-10), :415-455 (PSD batch == scalar, atol 1e-12).
(3) Hand-evaluated known answers of SURVEY.md Appendix B.
CPU-only; no CUDA.
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import gc_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHAINS = sorted(os.path.basename(p)[len("ref_chain_"):-4] for p in glob.glob(os.path.join(GOLDEN, "ref_chain_*.npz")))


def _close(a, b, rtol=1e-12, atol=0.0, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol + rtol * s)


@pytest.fixture(scope="module")
def pieces():
    return np.load(os.path.join(GOLDEN, "ref_pieces.npz"))


@pytest.fixture(scope="module")
def ref_certs():
    with open(os.path.join(GOLDEN, "ref_certs.json")) as f:
        return json.load(f)


def test_goldens_present():
    assert len(CHAINS) >= 5


# ---------------------------------------------------------------- (1) pieces
def test_kappa_matches_reference(pieces):
    _close(O.kappa_from_resultant_batch(pieces["kappa_in"]), pieces["kappa_out"], rtol=1e-14)
    sc = np.array([O.kappa_scalar(r) for r in pieces["kappa_in"]])
    _close(sc, pieces["kappa_scalar_out"], rtol=1e-14)


@pytest.mark.parametrize("B", [16, 48, 256, 1024])
def test_atlas_matches_reference(pieces, B):
    _close(O.create_fibonacci_atlas(B), pieces[f"atlas_{B}"], rtol=0, atol=1e-15)


def test_so3_se3_match_reference(pieces):
    for o, R, lg in zip(pieces["so3_omega"], pieces["so3_exp"], pieces["so3_log_of_exp"]):
        _close(O.so3_exp(o), R, rtol=0, atol=1e-15)
        _close(O.so3_log(O.so3_exp(o)), lg, rtol=0, atol=1e-14)
    for x, T in zip(pieces["se3_xi"], pieces["se3_exp"]):
        _close(O.se3_exp(x), T, rtol=0, atol=1e-14)
    for a, b, c in zip(pieces["se3_a"], pieces["se3_b"], pieces["se3_compose"]):
        _close(O.se3_compose(a, b), c, rtol=0, atol=1e-14)


def test_psd_projection_matches_reference(pieces):
    for M, P, cert in zip(pieces["psd_in"], pieces["psd_out"], pieces["psd_cert"]):
        Po, co = O.domain_projection_psd(M, 1e-12)
        _close(Po, P, rtol=1e-13, atol=1e-24)
        _close(co[:4], cert[:4], rtol=1e-9, atol=1e-20)
        assert co[5] == cert[5]
    Pb, cb = O.domain_projection_psd_batch(pieces["psd_in"], 1e-12)
    _close(Pb, pieces["psd_out"], rtol=1e-13, atol=1e-24)


def test_inv_mass_and_window_weights(pieces):
    inv, ratio = O.inv_mass(pieces["invmass_in"], 1e-12)
    _close(inv, pieces["invmass_inv"], rtol=1e-15)
    _close(ratio, pieces["invmass_ratio"], rtol=1e-15)
    _close(O.smooth_window_weights(pieces["sww_t"], 1.7e9, 1.7e9 + 0.1, 0.01), pieces["sww_w"], rtol=1e-14, atol=1e-300)


# ---------------------------------------------------------------- (1) whole chain
@pytest.mark.parametrize("name", CHAINS)
def test_chain_matches_reference(name, ref_certs):
    g = np.load(os.path.join(GOLDEN, f"ref_chain_{name}.npz"))
    cap, tau = int(g["in_n_points_cap"]), float(g["in_tau"])
    lam = g["in_point_lambda"] if "in_point_lambda" in g.files else None
    pb = O.point_budget_resample(g["in_points"], g["in_timestamps"], g["in_weights"], g["in_ring"], g["in_tag"], cap)
    # integer / byte outputs: bit exact
    assert np.array_equal(pb["ring"], g["pb_ring"]) and np.array_equal(pb["tag"], g["pb_tag"])
    assert np.array_equal(pb["points"], g["pb_points"]) and np.array_equal(pb["timestamps"], g["pb_timestamps"])
    assert [pb["n_input"], pb["n_output"]] == [int(g["pb_scalars"][0]), int(g["pb_scalars"][1])]
    _close(pb["weights"], g["pb_weights"], rtol=1e-15)
    cj = ref_certs[name]
    _close(pb["ess"], cj["point_budget"]["cert"]["support"]["ess_total"], rtol=1e-13)
    _close(pb["cert_support_frac"], cj["point_budget"]["cert"]["support"]["support_frac"], rtol=1e-15)
    _close(pb["cert_mass_epsilon_ratio"], cj["point_budget"]["cert"]["influence"]["mass_epsilon_ratio"], rtol=1e-15)

    dk = O.deskew_constant_twist(pb["points"], pb["timestamps"], pb["weights"], float(g["in_scan_start_time"]),
                                 float(g["in_scan_end_time"]), g["in_xi_body"])
    n_sel = pb["n_output"]
    # valid rows to 1e-13; padded rows carry epoch-scale twists (huge, finite): compare loosely
    _close(dk["points"][:n_sel], g["dk_points"][:n_sel], rtol=1e-13)
    if n_sel < cap:
        _close(dk["points"][n_sel:], g["dk_points"][n_sel:], rtol=1e-6)
    _close(dk["weights"], g["dk_weights"], rtol=1e-13, atol=1e-300)
    _close(dk["cert_support_frac"], cj["deskew"]["cert"]["support"]["support_frac"], rtol=1e-13)

    dirs = O.ray_directions(dk["points"], g["in_lidar_origin"])
    _close(dirs[:n_sel], g["dirs"][:n_sel], rtol=0, atol=1e-14)
    sa = O.bin_soft_assign(g["dirs"], g["atlas_dirs"], tau)      # same directions as the reference saw
    _close(sa["responsibilities"][:64], g["sa_resp_head"], rtol=1e-13, atol=1e-300)
    _close(sa["responsibilities"].sum(0), g["sa_resp_colsum"], rtol=1e-13)
    _close(sa["avg_entropy"], cj["soft_assign"]["effect"]["predicted"], rtol=1e-13)
    _close(sa["cert_ess_total"], cj["soft_assign"]["cert"]["support"]["ess_total"], rtol=1e-13)
    _close(sa["cert_support_frac"], cj["soft_assign"]["cert"]["support"]["support_frac"], rtol=1e-13)

    mm = O.scan_bin_moment_match(g["dk_points"], None, g["dk_weights"], sa["responsibilities"], lam, g["in_lidar_origin"])
    for k in ("N", "s_dir", "S_dir_scatter", "p_bar", "kappa_scan"):
        _close(mm[k], g["mm_" + k], rtol=1e-12)
    _close(mm["Sigma_p"], g["mm_Sigma_p"], rtol=1e-10)
    mc = cj["moment_match"]["cert"]
    _close(mm["ess"], mc["support"]["ess_total"], rtol=1e-12)
    _close(mm["support_frac"], mc["support"]["support_frac"], rtol=1e-12)
    _close(mm["max_eps_ratio_N"], mc["influence"]["mass_epsilon_ratio"], rtol=1e-12)
    # when no eigenvalue is floored this number is pure round-off of V diag V^T (differs between LAPACK
    # drivers): parity is absolute, at 1e-13 of the summed matrix norms
    psd_scale = float(np.sum(np.linalg.norm(g["mm_Sigma_p"], axis=(1, 2))))
    _close(mm["psd_projection_delta_total"], mc["influence"]["psd_projection_delta"], rtol=1e-6, atol=1e-13 * psd_scale)

    # chunked/fused formulation == materialised formulation
    fz = O.fused_scan_raw_sums(g["dk_points"], g["dk_weights"], g["atlas_dirs"], tau, g["in_lidar_origin"], lam, chunk=100)
    for k in ("N", "s_dir", "S_dir_scatter", "sum_p", "sum_ppT"):
        _close(fz[k], mm["raw"][k], rtol=1e-13)

    ms = {k[4:]: g[k] for k in g.files if k.startswith("map_") and k not in ("map_Rstar", "map_tstar")}
    md = O.compute_map_derived_stats(ms["S_dir"], ms["N_dir"], ms["N_pos"], ms["sum_p"], ms["sum_ppT"])
    _close(md["mu_dir"], g["md_mu_dir"], rtol=1e-14)
    _close(md["kappa"], g["md_kappa"], rtol=1e-12)
    _close(md["centroid"], g["md_centroid"], rtol=1e-14)
    _close(md["Sigma_c"], g["md_Sigma_c"], rtol=1e-11)
    _close(O.apply_forgetting(ms, 0.99)["sum_ppT"], g["forget_sum_ppT"], rtol=1e-15)

    for h, pose in enumerate(g["in_pose_preds"]):
        pw = O.world_pose(pose, np.zeros(6))
        _close(pw, g["pose_world"][h], rtol=0, atol=1e-14)
        mf = O.matrix_fisher_rotation_evidence(pw, g["mm_s_dir"], g["mm_S_dir_scatter"], g["mm_N"], ms["S_dir"],
                                               ms["S_dir_scatter"], ms["N_dir"])
        _close(mf["R_mf"], g["hyp_R_mf"][h], rtol=0, atol=1e-13)
        _close(mf["svd_singular_values"], g["hyp_svd_s"][h], rtol=1e-13)
        _close(mf["L_rot"], g["hyp_L_rot"][h], rtol=1e-12)
        _close(mf["delta_rot"], g["hyp_delta_rot"][h], rtol=0, atol=1e-13)
        _close(mf["h_rot"], g["hyp_h_rot"][h], rtol=1e-11)
        _close(mf["scan_scatter_metrics"]["eigenvalues"], g["hyp_scan_scatter_eigs"][h], rtol=1e-13)
        _close(mf["map_scatter_metrics"]["eigenvalues"], g["hyp_map_scatter_eigs"][h], rtol=1e-13)
        c = cj["matrix_fisher"][h]
        _close(mf["rot_nll"], c["effect"]["predicted"], rtol=1e-9, atol=1e-22)
        _close(mf["cert_cond"], c["cert"]["conditioning"]["cond"], rtol=1e-12)
        _close(mf["cert_nll_per_ess"], c["cert"]["mismatch"]["nll_per_ess"], rtol=1e-9, atol=1e-24)
        _close(mf["cert_directional_score"], c["cert"]["mismatch"]["directional_score"], rtol=1e-13)
        for k in ("linearity", "planarity", "sphericity", "anisotropy", "effective_rank"):
            _close(mf["scan_scatter_metrics"][k], c["scan_metrics"][k], rtol=1e-12, atol=1e-15)
        pt = O.planar_translation_evidence(pw, g["mm_p_bar"], g["mm_Sigma_p"], g["mm_N"], g["md_centroid"], g["md_Sigma_c"],
                                           ms["N_pos"], ms["S_dir_scatter"], ms["N_dir"], g["hyp_R_mf"][h])
        _close(pt["t_wls"], g["hyp_t_wls"][h], rtol=1e-11)
        _close(pt["L_trans"], g["hyp_L_trans"][h], rtol=1e-11)
        _close(pt["h_trans"], g["hyp_h_trans"][h], rtol=1e-9)
        _close(pt["delta_trans"], g["hyp_delta_trans"][h], rtol=1e-10)
        c = cj["planar_translation"][h]
        _close(pt["trans_nll"], c["effect"]["predicted"], rtol=1e-9, atol=1e-22)
        _close(pt["cert_eig_min"], c["cert"]["conditioning"]["eig_min"], rtol=1e-9)
        L22, h22 = O.build_combined_lidar_evidence_22d(mf["L_rot"], mf["h_rot"], pt["L_trans"], pt["h_trans"])
        _close(L22, g["hyp_L22"][h], rtol=1e-11)
        _close(h22, g["hyp_h22"][h], rtol=1e-9)
        assert np.count_nonzero(L22[6:, :]) == 0 and np.count_nonzero(L22[:, 6:]) == 0


def test_fused_chain_equals_staged_chain():
    g = np.load(os.path.join(GOLDEN, "ref_chain_stride3_pad.npz"))
    ms = {k[4:]: g[k] for k in g.files if k.startswith("map_") and k not in ("map_Rstar", "map_tstar")}
    kw = dict(n_points_cap=int(g["in_n_points_cap"]), scan_start_time=float(g["in_scan_start_time"]),
              scan_end_time=float(g["in_scan_end_time"]), xi_body=g["in_xi_body"], lidar_origin=g["in_lidar_origin"],
              bin_directions=g["atlas_dirs"], tau=float(g["in_tau"]), map_stats=ms, pose_pred=g["pose_world"][1])
    a = O.lidar_scan_evidence(g["in_points"], g["in_timestamps"], g["in_weights"], g["in_ring"], g["in_tag"], materialize=True, **kw)
    b = O.lidar_scan_evidence(g["in_points"], g["in_timestamps"], g["in_weights"], g["in_ring"], g["in_tag"], materialize=False, **kw)
    _close(b["L_lidar"], a["L_lidar"], rtol=1e-10)
    _close(b["h_lidar"], a["h_lidar"], rtol=1e-9)
    _close(a["L_lidar"], g["hyp_L22"][1], rtol=1e-10)
    _close(b["soft_assign"]["avg_entropy"], a["soft_assign"]["avg_entropy"], rtol=1e-12)


# ---------------------------------------------------------------- (2) the reference's own numeric tests
def test_ref_kappa_batch_matches_scalar():            # test_audit_invariants.py:398-413
    Rv = np.array([0.1, 0.3, 0.5, 0.7, 0.85])
    kb = O.kappa_from_resultant_batch(Rv)
    for i, R in enumerate(Rv):
        np.testing.assert_allclose(kb[i], O.kappa_scalar(float(R)), rtol=1e-10)


def test_ref_kappa_smooth():                           # test_audit_invariants.py:103-119
    k = O.kappa_from_resultant_batch(np.linspace(0.01, 0.99, 100))
    dk = np.diff(k)
    assert np.max(np.abs(dk)) < 100 * np.median(np.abs(dk))


@pytest.mark.parametrize("omega,atol", [([0.01, -0.02, 0.015], 1e-10), ([0.5, -0.7, 0.3], 1e-10),
                                        ([1.5, -1.2, 0.8], 1e-9), ([np.pi - 0.01, 0.0, 0.0], 1e-8)])
def test_ref_so3_roundtrip(omega, atol):               # test_audit_invariants.py:224-268
    R = O.so3_exp(np.array(omega))
    np.testing.assert_allclose(O.so3_exp(O.so3_log(R)), R, atol=atol)


def test_ref_so3_log_exp_random():                     # test_audit_invariants.py:270-284
    rng = np.random.default_rng(456)
    for _ in range(10):
        om = rng.normal(size=3) * 0.5
        np.testing.assert_allclose(O.so3_log(O.so3_exp(om)), om, atol=1e-9)


def test_ref_psd_extreme_negative():                   # test_audit_invariants.py:121-140
    P, cert = O.domain_projection_psd(np.diag([1.0, -1000.0, 1.0]), 1e-12)
    assert np.all(np.linalg.eigvalsh(P) > 0) and np.isfinite(P).all()


def test_ref_psd_batch_matches_scalar():               # test_audit_invariants.py:415-455
    rng = np.random.default_rng(111)
    mats = np.stack([(lambda A: A @ A.T - 0.1 * np.eye(3))(rng.normal(size=(3, 3))) for _ in range(5)])
    Pb, _ = O.domain_projection_psd_batch(mats, 1e-6)
    for i in range(5):
        np.testing.assert_allclose(Pb[i], O.domain_projection_psd(mats[i], 1e-6)[0], atol=1e-12)


def test_ref_structural_contracts():                   # archive/legacy_tests/test_operators.py:47-73,137-151
    rng = np.random.default_rng(42)
    pb = O.point_budget_resample(rng.normal(size=(10000, 3)), np.linspace(0, 1, 10000), np.ones(10000))
    assert pb["points"].shape[0] <= O.GC_N_POINTS_CAP
    assert abs(pb["total_mass_out"] - pb["total_mass_in"]) < 1e-6
    assert abs(pb["weights"].sum() - pb["total_mass_in"]) < 1e-6
    d = rng.normal(size=(50, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    sa = O.bin_soft_assign(d, O.create_fibonacci_atlas(20), 0.1)
    np.testing.assert_allclose(sa["responsibilities"].sum(1), 1.0, atol=1e-6)


# ---------------------------------------------------------------- (3) SURVEY.md Appendix B known answers
def test_appendix_b_known_answers():
    k = O.kappa_from_resultant_batch(np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.8, 0.85, 0.9, 0.99, 1.0]))
    np.testing.assert_allclose(k, [0.0, 0.3020198969281307, 0.9593395551421199, 1.8332607197301192, 3.3496179256527054,
                                   3.13304056205783, 2.1865615031302803, 1.9608486972242263, 4.088241621349586,
                                   860.4783517808435], rtol=1e-12)
    a = O.create_fibonacci_atlas(48)
    np.testing.assert_allclose(a[0], [0.0735831985, -0.1892568408, 0.9791666667], atol=1e-9)
    np.testing.assert_allclose(a[47], [0.1260769584, -0.1591767554, -0.9791666667], atol=1e-9)
    for (n, cap), (stride, nsel) in {(100, 8192): (1, 100), (10000, 8192): (2, 5000), (30000, 8192): (4, 7500),
                                     (30000, 32768): (1, 30000), (1048576, 8192): (128, 8192)}.items():
        pb = O.point_budget_resample(np.zeros((n, 3)), np.zeros(n), np.ones(n), n_points_cap=cap)
        assert (pb["stride"], pb["n_output"]) == (stride, nsel)
        assert np.array_equal(pb["indices"], np.arange(0, n, max(1, int(np.ceil(n / cap)))))


# ---------------------------------------------------------------- IMU preintegration -> deskew twist
@pytest.fixture(scope="module")
def imu_ref():
    return np.load(os.path.join(GOLDEN, "ref_imu.npz"))


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_imu_preintegration_matches_reference(imu_ref, case):
    z = imu_ref
    g = lambda k: z[f"{case}_{k}"]
    out = O.imu_deskew_twist(g("stamps"), g("gyro"), g("accel"), float(g("t0")), float(g("t1")), float(g("sigma")), g("rotvec0"),
                             g("bg"), g("ba"), g("grav"))
    np.testing.assert_allclose(out["weights"], g("w"), rtol=1e-13, atol=1e-300)
    for k in ("delta_pose", "delta_R", "delta_p", "delta_v", "a_body_mean", "a_world_nog_mean", "a_world_mean"):
        np.testing.assert_allclose(out[k], g(k), rtol=1e-11, atol=1e-13, err_msg=k)
    assert abs(out["ess"] / float(g("ess")) - 1) < 1e-13 and abs(out["dt_sum"] / float(g("dt_sum")) - 1) < 1e-13
    np.testing.assert_allclose(out["xi_body"], g("xi"), rtol=1e-10, atol=1e-13)


def test_se3_log_matches_reference(imu_ref):
    for p, ref in zip(imu_ref["se3log_in"], imu_ref["se3log_out"]):
        np.testing.assert_allclose(O.se3_log(p), ref, rtol=1e-10, atol=1e-13)


# ---------------------------------------------------------------- PointCloud2 ingest
PC2_FIELD_NAMES = {0: "x", 1: "y", 2: "z", 3: "ring", 4: "time", 5: "intensity"}


def pc2_case(z, name):
    fields = [(PC2_FIELD_NAMES[int(k)], int(off), int(dt)) for k, off, dt in z[f"{name}_fields"]]
    return z[f"{name}_raw"], fields, int(z[f"{name}_step"]), int(z[f"{name}_n"]), tuple(int(v) for v in z[f"{name}_stamp"])


@pytest.mark.parametrize("name", ["kimera", "ns_u32", "no_time", "f64"])
def test_pointcloud2_parse_matches_reference(name):
    z = np.load(os.path.join(GOLDEN, "ref_pointcloud2.npz"))
    raw, fields, step, n, stamp = pc2_case(z, name)
    out = O.parse_pointcloud2_vlp16(raw, fields, step, n, stamp)
    assert np.array_equal(out["points"], z[f"{name}_points_lidar"])          # byte work: bit-exact
    assert np.array_equal(out["timestamps"], z[f"{name}_t"])
    assert np.array_equal(out["ring"], z[f"{name}_ring"]) and np.array_equal(out["tag"], z[f"{name}_tag"])
    np.testing.assert_allclose(out["weights"], z[f"{name}_w"], rtol=1e-15, atol=0)
    np.testing.assert_allclose(O.lidar_points_to_base(out["points"], z["R_base_lidar"], z["t_base_lidar"]), z[f"{name}_points_base"],
                               rtol=1e-15, atol=1e-15)


def test_scan_time_bounds_against_reference_source():
    """O.scan_time_bounds against the reference's own statements (backend_node.py:1776-1786), cut out of the source as text and
    executed by tests/golden/generate_goldens.py::run_time_bounds: bit-exact on every case (spread, no spread, spread below the
    1e-9 threshold; header stamp before / inside / after the scan; epoch-scale and small time bases)."""
    z = np.load(os.path.join(GOLDEN, "ref_time_bounds.npz"))
    for i in range(int(z["n_cases"])):
        stamp, t0, t1 = z[f"io_{i}"]
        assert O.scan_time_bounds(z[f"ts_{i}"], float(stamp)) == (float(t0), float(t1)), i
