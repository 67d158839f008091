#!/usr/bin/env python
"""Fill the table of BASELINE.md section 4 from the bench lines kept under profiles/ (r2_bench_n{1,2,4,8}.json = `bench.py --gpus N`
under torchrun, r2_bench_cfg5_n8_{fp64,fp32c}.json = configs[4] through the same code, r2_bench_reference.json = the CPU arm).

    python tools/results_table.py > profiles/r2_results_table.md
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    f = os.path.join(P, name)
    if not os.path.exists(f):
        return None
    return json.loads(open(f).read().strip().splitlines()[-1])


def f3(x):
    return "%.3g" % x


runs = {n: load(f"r2_bench_n{n}.json") for n in (1, 2, 4, 8)}
runs = {n: d for n, d in runs.items() if d}
ns = sorted(runs)
ref = load("r2_bench_reference.json")
one = runs[1]
print("# round 2: the table of BASELINE.md section 4, filled in from driver-comparable runs\n")
print("Every figure below comes from `python bench.py --gpus N` (N > 1: under `torch.distributed.run`, one rank per GPU of one box), the command the")
print("driver runs, on the code of the commit that holds this file; the lines themselves are `profiles/r2_bench_n{N}.json`. SM clock during the")
print("timed regions: " + ", ".join(f"N={n}: {runs[n]['clocks']['sm_mhz']:.0f} MHz {runs[n]['clocks']['reasons']}" for n in ns) + ".")
print("Time per scan = CUDA events around the timed region, max over ranks; `scan_moments` alone = events on its own stream.\n")

print("## configs[3] (BASELINE.md cfg 4): 1 048 576 points x 256 bins, sharded by points, one fused NVLink exchange per scan (strong scaling)\n")
print("| GPUs | precision | whole scan ms (median / best) | points/s | scan_moments ms | exchange + finalise + evidence (not hidden) ms | efficiency vs 1 GPU | e2e points/s (host arrays, H2D + D2H inside) | sharded vs unsharded max rel | ranks bit-identical | FMA-issue fraction |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for prec in ("fp64", "fp32c"):
    base = None
    for n in ns:
        d = runs[n]
        if prec == "fp64":
            det, par, e2e, val, cf = d["details"], d["parity"], d["e2e"]["value"], d["value"], d["roofline"]["compute_frac"]
        else:
            x = d["extra"]["cfg4_1Mx256_fp32c"]
            det, par, e2e, val, cf = x, x["parity"], x["e2e"]["value"], x["points_per_sec"], x["roofline"].get("compute_frac", float("nan"))
        ms = det["ms_per_scan"]
        base = base or ms
        print(f"| {n} | {prec} | {ms:.4f} ({det['ms_per_scan_median']:.4f} / {det['ms_per_scan_best']:.4f}) | {val:.3e} | {det['scan_moments_ms']:.4f} | "
              f"{det['rest_of_scan_ms']:.4f} | {base / (n * ms):.2f} | {e2e:.3e} | {max(par['vs_unsharded_max_rel'].values()):.1e} | "
              f"{par['ranks_bit_identical']} | {cf:.2f} |")
if any("cfg4_1Mx256_nccl_all_gather_route" in runs[n]["extra"] for n in ns):
    print("\nThe round-1 route (two launches + `all_gather_into_tensor` + three launches), same runs: " + ", ".join(
        f"N={n}: {runs[n]['extra']['cfg4_1Mx256_nccl_all_gather_route']['ms_per_scan']:.3f} ms" for n in ns if "cfg4_1Mx256_nccl_all_gather_route" in runs[n]["extra"]) + " per fp64 scan.")

print("\n## configs[2] (cfg 3): 64 pose hypotheses x 30 000 points (cap 32768) x 48 bins, sharded by hypotheses, records gathered by one peer-exchange launch\n")
print("| GPUs | precision | 64-hypothesis batch ms (pipelined / one at a time) | hypothesis-scans/s | efficiency vs 1 GPU | batch vs single-hypothesis max rel | gathered records bit-identical |")
print("|---|---|---|---|---|---|---|")
for key, prec in (("cfg3_64hyp", "fp64"), ("cfg3_64hyp_fp32c", "fp32c")):
    base = None
    for n in ns:
        x = runs[n]["extra"][key]
        ms = x["ms_per_batch"]
        base = base or ms
        print(f"| {n} | {prec} | {ms:.4f} / {x['ms_per_batch_one_at_a_time']:.4f} | {x['hypothesis_scans_per_sec']:.3e} | {base / (n * ms):.2f} | "
              f"{x['parity']['batch_vs_single_hypothesis_max_rel']:.1e} | {x['parity']['gathered_bit_identical_across_ranks']} |")

print("\n## configs[1] (cfg 2): stream of 200 scans x ~30 000 points (cap 8192: stride 4, 7500 rows), 48 bins, map + IW carried -- replicas only\n")
s = one["extra"]["cfg2_stream"]
e = s["e2e"]
print(f"* resident, consecutive 200-scan batches pipelined: {s['ms_per_step']:.3f} ms per batch = {s['points_per_sec']:.3e} raw points/s "
      f"({s['processed_rows_per_sec']:.3e} processed rows/s), {s['scans_per_sec']:.3e} scans/s; one batch at a time {s['ms_per_step_one_batch_at_a_time']:.3f} ms; "
      f"`scan_moments` {s['roofline'].get('ms_per_launch', float('nan')):.3f} ms per launch, FMA-issue fraction {s['roofline'].get('compute_frac', s['roofline'].get('compute_frac_at_1965MHz', float('nan'))):.2f}, "
      f"HBM fraction {s['roofline']['frac']:.3f}.")
print(f"* two-stage host calls (host arrays in, records out, every scan): {e['us_per_scan']:.1f} us per scan = {e['value']:.3e} points/s; blocking call "
      f"{e['blocking_call']['us_per_scan']:.1f} us (p50 {e['blocking_call']['latency_us']['p50']:.0f}, p99 {e['blocking_call']['latency_us']['p99']:.0f}); "
      f"from the raw PointCloud2 message + IMU buffer {e['from_message']['us_per_scan']:.1f} us per scan.")
print("* N replicas (the stream does not shard): " + ", ".join(f"N={n}: {runs[n]['extra']['cfg2_stream']['points_per_sec']:.3e} points/s" for n in ns) + ".")

print("\n## configs[4] (cfg 5): 16 777 216 points x 1024 bins through the same sharded code at 8 GPUs\n")
print("| GPUs | precision | whole scan ms | points/s | pairs/s | scan_moments ms | sharded vs unsharded max rel | ranks bit-identical |")
print("|---|---|---|---|---|---|---|---|")
for prec in ("fp64", "fp32c"):
    d = load(f"r2_bench_cfg5_n8_{prec}.json")
    if d:
        det, par = d["details"], d["parity"]
        print(f"| 8 | {prec} | {det['ms_per_scan']:.3f} | {d['value']:.3e} | {d['pairs_per_sec']:.3e} | {det['scan_moments_ms']:.3f} | "
              f"{max(par['vs_unsharded_max_rel'].values()):.1e} | {par['ranks_bit_identical']} |")
print("\n(1 GPU, round 1 sweep `profiles/r1_cfg5_sweep_16Mx1024.md`: fp64 64.7 ms, fp32c 26.1 ms for `scan_moments`.)")

print("\n## CPU arm (`bench.py --impl reference`, the NumPy restatement; JAX is not installed) and roofline of the headline\n")
if ref:
    cb = ref["cpu_baseline"]
    print(f"* reference arm: {ref['value']:.3e} points/s on {cb['cores']} threads (single thread {cb['single_thread_value']:.3e}); sample: {cb['sample']}; "
          f"native libraries of this repo mapped by that process: {ref.get('loaded_native_libraries')}.")
r = one["roofline"]
print(f"* headline at 1 GPU: value {one['value']:.3e} points/s resident, e2e {one['e2e']['value']:.3e} points/s; ratio to the CPU arm "
      f"{one['e2e']['value'] / ref['value']:.0f}x (e2e) / {one['value'] / ref['value']:.0f}x (resident)." if ref else "")
print(f"* `{r['kernel']}`: {r['ms_per_launch']:.4f} ms per launch, {r['achieved']:.1f} GB/s algorithmic = {r['frac']:.4f} of the measured HBM peak "
      f"({r['peak']} GB/s), DRAM traffic {r['traffic'] / r['algorithmic_bytes_per_launch']:.3f}x algorithmic; bound by FP64 issue: "
      f"{r['pairs_per_sec']:.3e} pairs/s = {r['compute_frac']:.2f} of the FMA-issue ceiling ({r['compute_ceiling_pairs_per_sec']:.3e}).")
print(f"* GPU vs the oracle on the CPU sample (max rel): {one['parity']['vs_oracle_on_cpu_sample']}.")
