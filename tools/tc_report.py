#!/usr/bin/env python
"""fp32c scan_moments: tensor-core kernel (gcl_scan_tc.cu) against the CUDA-core kernel (GCL_K1_TC=0), time and error.

    python tools/tc_report.py            # runs itself twice (the switch is read once per process) and prints a table

Error = max |fp32c - fp64| / max |fp64| per raw-sum array, where fp64 is the fp64 kernel on the same points (itself <= 1e-13 of the
oracle, tests/test_gpu_parity.py).  Time = best of 3 x 20 launches of scan_moments alone (CUDA events), inputs rotated through copies.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure():
    import numpy as np
    import torch
    import fl_slam_b200 as g
    tau = g.synth.DEFAULT_TAU
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for name, n, cap, bins, n_scans in (("1Mx256", 1 << 20, 1 << 20, 256, 1), ("1Mx128", 1 << 20, 1 << 20, 128, 1), ("stream200x30k_48", 30000, 8192, 48, 200),
                                        ("64hyp_30k_48", 30000, 32768, 48, 64)):
        scans = [g.synth.make_scan(n + 37 * (k % 11), seed=2000 + k, dense=n > 100000) for k in range(n_scans if name.startswith("stream") else 1)]
        res = {}
        for precision in ("fp64", "fp32c"):
            pl = g.LidarEvidencePipeline(n_bins=bins, tau=tau, precision=precision, n_points_cap=cap, lidar_origin=scans[0].lidar_origin)
            n_copies = 6 if n > 100000 else 1
            ups = [[pl.upload_scan(s.points, s.timestamps, s.weights, s.scan_start_time, s.scan_end_time) for s in scans] for _ in range(n_copies)]
            if name.startswith("64hyp"):
                rng = np.random.default_rng(3)
                xis = scans[0].xi_body[None] * (1 + 0.02 * rng.normal(size=(64, 6)))
                jobs = [pl.jobs_to_device([pl.make_job(u[0], xis[h]) for h in range(64)]) for u in ups]
                nj = 64
            else:
                jobs = [pl.jobs_to_device([pl.make_job(u[k], scans[k].xi_body) for k in range(len(scans))]) for u in ups]
                nj = len(scans)
            max_sel = max(pl.selected_rows(s.points.shape[0], cap)[1] for s in scans)
            ctas = pl.plan_ctas(nj, max_sel)
            b = pl._buffers(nj, ctas)
            i = [0]

            def fn():
                pl.launch_moments(jobs[i[0] % n_copies], None, nj, ctas, ups[0][0].layout, b["partials"])
                i[0] += 1
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                ev0.record()
                for _ in range(20):
                    fn()
                ev1.record()
                torch.cuda.synchronize()
                best = min(best, ev0.elapsed_time(ev1) / 20)
            i[0] = 0
            fn()
            pl.launch_finalize(b["partials"], ctas, pl.rec, pl.rec * ctas, jobs[0], None, nj, True, b["reduced"], b["stats"])
            torch.cuda.synchronize()
            res[precision] = (best, b["reduced"].detach().cpu().numpy().copy(), b["stats"].detach().cpu().numpy().copy())
        B = bins
        r64, r32 = res["fp64"][1], res["fp32c"][1]
        names = (("N", 0, 1), ("s_dir", 1, 4), ("S_dir_scatter", 4, 10), ("sum_p", 10, 13), ("sum_ppT", 13, 19))
        err = {}
        for nm, a, e in names:
            x64, x32 = r64[:, a * B:e * B], r32[:, a * B:e * B]
            err[nm] = float(np.max(np.abs(x32 - x64)) / np.max(np.abs(x64)))
        s64, s32 = res["fp64"][2], res["fp32c"][2]
        for nm, a, e in (("p_bar", 13, 16), ("Sigma_p", 16, 25), ("kappa", 25, 26)):
            x64, x32 = s64[:, a * B:e * B], s32[:, a * B:e * B]
            err[nm] = float(np.max(np.abs(x32 - x64)) / np.max(np.abs(x64)))
        pairs = sum(pl.selected_rows(s.points.shape[0], cap)[1] for s in scans) * bins * (64 if name.startswith("64hyp") else 1)
        out[name] = {"fp32c_us": res["fp32c"][0] * 1e3, "fp64_us": res["fp64"][0] * 1e3, "fp32c_Gpairs_per_s": pairs / res["fp32c"][0] * 1e-6,
                     "max_rel_err_vs_fp64": err}
    print("TCREPORT " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        measure()
        sys.exit(0)
    rows = {}
    for label, val in (("tensor-core kernel (GCL_K1_TC=2: every bin count)", "2"), ("CUDA-core kernel (GCL_K1_TC=0)", "0")):
        env = dict(os.environ, GCL_K1_TC=val)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=900)
        line = [l for l in p.stdout.splitlines() if l.startswith("TCREPORT ")]
        if not line:
            print(label, "FAILED\n", p.stdout[-2000:], p.stderr[-3000:])
            continue
        rows[label] = json.loads(line[0][9:])
    print(json.dumps(rows, indent=1))
