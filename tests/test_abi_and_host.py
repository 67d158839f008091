"""CPU-only checks: the C-ABI library loads and exports every symbol include/gcl_b200.h declares (no compute calls),
the host-side mirror of the reference interface (certificates, constants, manifest, planning helpers) behaves as the
reference does, and the product never routes through the oracle."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def g():
    import fl_slam_b200
    return fl_slam_b200


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gcl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(g):
    lib = ctypes.CDLL(g._abi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libgcl_b200.so does not export {n}"
    assert set(g._abi.EXPORTED) == set(names), "the ctypes table and the header disagree"
    assert lib.gcl_abi_version() == 1


def test_record_sizes_match_header_constants(g):
    lib = g._abi.lib
    for B in (1, 7, 48, 256, 1024):
        assert lib.gcl_partial_record_doubles(B) == 19 * B + 8
        assert lib.gcl_stats_record_doubles(B) == 38 * B + 24
    assert g._abi.EV["L22"] + 484 <= g._abi.GCL_EVID_DOUBLES
    assert ctypes.sizeof(g._abi.GclJob) == 120


def test_fails_loudly_without_a_gpu(g):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g._abi.GclError):
        g._abi.handle(0)
    with pytest.raises(g._abi.GclError):
        g.LidarEvidencePipeline(n_bins=48, tau=0.1)
    with pytest.raises(g._abi.GclError):
        g.point_budget_resample(np.zeros((4, 3)), np.zeros(4), np.ones(4))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "fl-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "gc_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_constants_match_reference_values(g):
    C = g.constants   # values asserted by the reference's own tests: test_geometric_compositional_invariants.py:117-143
    assert C.GC_N_POINTS_CAP == 8192 and C.GC_CHART_ID == "GC-RIGHT-01" and C.GC_K_HYP == 4
    assert 1e-15 < C.GC_EPS_PSD < 1e-6 and 1e-15 < C.GC_EPS_MASS < 1e-6 and 1e-12 < C.GC_EPS_LIFT < 1e-6
    assert (C.GC_EPS_R, C.GC_KAPPA_BLEND_R0, C.GC_KAPPA_BLEND_TAU) == (1e-6, 0.8, 0.03)
    assert (C.GC_TIME_WARP_SIGMA_FRAC, C.GC_WEIGHT_FLOOR) == (0.1, 1e-12)


def test_certificates_serialise_like_the_reference(g):
    """to_dict() of our CertBundle has exactly the keys of the reference's (golden written by the reference's own
    certificates.py through generate_goldens.py), and reproduces its numbers when fed the same fields."""
    ref = json.load(open(os.path.join(GOLDEN, "ref_certs.json")))["stride3_pad"]["point_budget"]["full_to_dict"]
    cert = g.CertBundle.create_approx(
        chart_id=ref["chart_id"], anchor_id=ref["anchor_id"], triggers=list(ref["approximation_triggers"]),
        support=g.SupportCert(**ref["support"]),
        influence=g.InfluenceCert.identity().with_overrides(mass_epsilon_ratio=ref["influence"]["mass_epsilon_ratio"]))
    ours = cert.to_dict()
    assert set(ours) == set(ref)
    for k in ("conditioning", "support", "mismatch", "excitation", "influence", "overconfidence", "compute"):
        assert set(ours[k]) == set(ref[k]), k
    assert ours["support"] == ref["support"] and ours["influence"] == ref["influence"]
    assert ours["total_trigger_magnitude"] == ref["total_trigger_magnitude"]
    assert ours["exact"] is False and ours["frobenius_applied"] is False
    ex = g.CertBundle.create_exact(chart_id="GC-RIGHT-01", anchor_id="a")
    assert ex.exact and ex.approximation_triggers == [] and ex.total_trigger_magnitude() == 0.0
    eff = g.ExpectedEffect("predicted_ess", 3.0).to_dict()
    assert eff == {"objective_name": "predicted_ess", "predicted": 3.0, "realized": None}


def test_selected_rows_follow_point_budget_rule(g):
    f = g.LidarEvidencePipeline.selected_rows   # point_budget.py:160,:70-71
    assert f(100, 8192) == (1, 100) and f(10000, 8192) == (2, 5000) and f(30000, 8192) == (4, 7500)
    assert f(30000, 32768) == (1, 30000) and f(1048576, 8192) == (128, 8192) and f(0, 8192) == (1, 0)
    for n in (1, 8191, 8192, 8193, 16384, 16385, 99999):
        s, k = f(n, 8192)
        assert k == len(range(0, n, s)) and k <= 8192


def test_synthetic_scans_are_deterministic(g):
    a, b = g.synth.make_scan(5000, seed=1234), g.synth.make_scan(5000, seed=1234)
    assert np.array_equal(a.points, b.points) and np.array_equal(a.weights, b.weights)
    assert a.points.dtype == np.float64 and a.ring.dtype == np.uint8 and a.timestamps.min() >= a.scan_start_time
    assert np.all(a.weights > 0) and np.all(np.isfinite(a.points))
    R, t = g.synth.known_map_transform(3)
    assert abs(np.linalg.det(R) - 1) < 1e-12 and abs(np.arccos((np.trace(R) - 1) / 2) - 0.05) < 1e-12


def test_manifest_entries(g):
    class P:  # the manifest only reads these attributes
        precision, n_points_cap, n_bins, tau, eps_psd, eps_mass = "fp64", 8192, 48, 0.1, 1e-12, 1e-12
    m = g.runtime_manifest_entries(P(), n_gpus=8, shard_mode="points")
    for k in ("point_budget", "deskew", "bin_soft_assign", "scan_bin_moment_match", "lidar_evidence", "domain_projection_psd", "se3"):
        assert "sm_100a" in m["backends"][k]
    assert m["B_BINS"] == 48 and m["tau_soft_assign"] == 0.1 and m["precision_mode"] == "fp64"
    assert m["n_gpus"] == 8 and m["shard_mode"] == "points" and m["pose_evidence_backend"] == "bins"


def test_launch_planning_without_a_gpu(g):
    """Host-side planning entry points need no device: CTAs per job fill whole waves of 148 SMs, workspaces grow with the
    problem."""
    lib = g._abi.lib
    assert lib.gcl_plan_ctas_per_job(None, 1, 1 << 20, 256, 1) == 148           # one big scan: one CTA per SM
    assert lib.gcl_plan_ctas_per_job(None, 1, 7500, 48, 0) == 79                # one small scan: 96-row CTAs
    c = lib.gcl_plan_ctas_per_job(None, 200, 7500, 48, 0)                       # a stream: waves x (rows per CTA + fixed cost) minimal
    waves = -(-200 * c // 148)
    cost = lambda k: -(-200 * k // 148) * (-(-7500 // k) + 128)
    assert cost(c) == min(cost(k) for k in range(1, 30)) and waves * 148 - 200 * c < 148
    assert lib.gcl_plan_ctas_per_job(None, 1, 0, 48, 0) >= 1
    w = [lib.gcl_scan_chain_workspace_doubles(s, h, b) for s, h, b in ((1, 1, 48), (200, 1, 48), (200, 4, 48), (200, 4, 256))]
    assert w == sorted(w) and w[0] >= 16 * 48


# ---- message path: host-side pieces (the device side is covered by the gpu tests) -------------------------------------------
@pytest.mark.parametrize("time_field,step", [("time_f32", 22), ("t_f64", 26), ("t_u32ns", 22), (None, 18)])
def test_synthetic_pointcloud2_parses_back_with_the_oracle(g, time_field, step):
    """synth.make_pointcloud2 writes the VLP-16 message layout the reference parser reads (backend_node.py:377-468): the oracle's
    restatement of that parser recovers points (to float32 rounding), ring, weights and the per-point times."""
    from oracle import gc_oracle as O
    sc = g.synth.make_scan(3000, seed=5, with_times=time_field is not None)
    R, t = g.synth.rotation_from_rotvec(np.array([0.02, -0.01, 0.3])), sc.lidar_origin
    base = sc.scan_start_time - (0.5 if time_field == "t_u32ns" else 12.25)
    m = g.synth.make_pointcloud2(sc, R, t, time_field=time_field, time_base=base, header_stamp_sec=12.0)
    assert m["point_step"] == step and m["data"].size == 3000 * step
    par = O.parse_pointcloud2_vlp16(m["data"], m["fields"], m["point_step"], m["n_points"], header_stamp=(12.0, 0))
    assert np.max(np.abs(O.lidar_points_to_base(par["points"], R, t) - sc.points)) < 5e-6
    assert np.array_equal(par["ring"], sc.ring)
    np.testing.assert_allclose(par["weights"], sc.weights, rtol=0, atol=2e-6)   # the window sees float32 coordinates
    if time_field is None:
        assert np.all(par["timestamps"] == 12.0)
    else:
        np.testing.assert_allclose(par["timestamps"], sc.timestamps - base, rtol=0, atol=2e-6)


def test_scan_time_bounds_rule():
    """pkg/backend/backend_node.py:1776-1786."""
    from oracle import gc_oracle as O
    assert O.scan_time_bounds(np.array([5.0, 5.0, 5.0]), 7.0) == (5.0, 7.0)                 # no per-point spread: [min, stamp]
    assert O.scan_time_bounds(np.array([5.0, 5.0 + 5e-10]), 4.0) == (5.0, 4.0)              # below the 1e-9 threshold, even backwards
    assert O.scan_time_bounds(np.array([5.0, 5.1]), 5.05) == (5.0, 5.1)                     # stamp inside
    assert O.scan_time_bounds(np.array([5.0, 5.1]), 4.9) == (4.9, 5.1)                      # stamp before
    assert O.scan_time_bounds(np.array([5.0, 5.1]), 5.3) == (5.0, 5.3)                      # stamp after


def test_message_submit_argument_checks_need_no_gpu(g):
    """The layout / argument checks of submit_scan_msg_host run before anything touches the device."""
    pl = object.__new__(g.LidarEvidencePipeline)
    pl._msg_layouts = {}
    f3 = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7)]
    with pytest.raises(ValueError):
        pl.submit_scan_msg_host(0, data=bytes(16), fields=f3 + [("ring", 12, 4)], point_step=16, n_points=1)          # neither twist nor IMU
    with pytest.raises(RuntimeError):
        pl.submit_scan_msg_host(0, data=bytes(12), fields=f3, point_step=12, n_points=1, xi_body=np.zeros((1, 6)))    # no ring field
    with pytest.raises(ValueError):
        pl.submit_scan_msg_host(0, data=bytes(16), fields=f3 + [("ring", 12, 9)], point_step=16, n_points=1, xi_body=np.zeros((1, 6)))
    with pytest.raises(ValueError):
        pl.submit_scan_msg_host(0, data=bytes(8), fields=f3 + [("ring", 12, 4)], point_step=16, n_points=1, xi_body=np.zeros((1, 6)))
    c = g.LidarEvidencePipeline._c64
    a = np.zeros((4, 3))
    assert c(a, 12) is a and c(a, 0, 3) is a
    assert c([[1, 2, 3]], 3).dtype == np.float64
    with pytest.raises(ValueError):
        c(a, 9)
    with pytest.raises(ValueError):
        c(np.zeros(7), 0, 3)


def test_numa_placement_helper_is_inert_without_nvml_or_with_one_node():
    """fl_slam_b200.hostmem.bind_to_gpu_numa_node (multi-rank bench runs): never raises, never widens the thread's CPU set, says why
    when it leaves the placement alone (this container: no NVML; the GPU pool: single-node VMs)."""
    import os
    from fl_slam_b200 import hostmem
    before = os.sched_getaffinity(0)
    info = hostmem.bind_to_gpu_numa_node(0)
    after = os.sched_getaffinity(0)
    assert set(info) >= {"applied", "numa_nodes", "cpus"} and info["numa_nodes"] >= 1
    assert after <= before and info["cpus"] == len(after)
    assert info["applied"] or "why" in info
    os.environ["GCL_NUMA_BIND"] = "0"
    try:
        assert hostmem.bind_to_gpu_numa_node(0)["why"] == "GCL_NUMA_BIND=0"
    finally:
        del os.environ["GCL_NUMA_BIND"]
