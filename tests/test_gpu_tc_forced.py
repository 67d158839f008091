"""The tensor-core fp32c kernel at EVERY bin count it supports (48-, 128- and 256-column variants).

By default the library routes fp32c scans to the tensor-core kernel only where it is the faster one (more than 128 bins, see
tc_path_supported in csrc/gcl_scan_tc.cu); the switch GCL_K1_TC is read once per process, so the fp32c parity cases are re-run here
in a child process with GCL_K1_TC=2 (tensor cores wherever the kernel can) and with GCL_K1_TC=0 (CUDA cores everywhere).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["2", "0"])
def test_fp32c_parity_with_the_kernel_choice_forced(mode):
    pytest.importorskip("torch")
    env = dict(os.environ, GCL_K1_TC=mode)
    sel = "fp32c and (fused_moments_vs_oracle or fused_pipeline_against_reference_goldens or bit_reproducible or degenerate_scans)"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-k", sel,
                        "-p", "no:cacheprovider"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("switch,mode", [("GCL_K1_WS", "1"), ("GCL_K1_WS", "2"), ("GCL_K1_WS", "3"), ("GCL_K1_MMA", "0"), ("GCL_K1_MMA", "1"), ("GCL_K1_MMA", "2"), ("GCL_K1_MMA", "3"), ("GCL_K1_MMA", "4")])
def test_fp64_parity_with_the_opt_in_kernels_forced(switch, mode):
    """GCL_K1_WS = 1 / 2 / 3: the warp-specialised fp64 kernel (129..256 bins; three splits of the warps between accumulating, phase-A and
    per-point roles); GCL_K1_MMA = 0 / 1 / 2 / 3 / 4: mma.sync.m8n8k4.f64 never / phase B with 8 warps / phase B with 16 warps / both phases
    (the default) / both phases with 128 threads and 32-point tiles (two CTAs per SM) -- same parity
    bars, bit-reproducible from run to run."""
    pytest.importorskip("torch")
    env = dict(os.environ, **{switch: mode})
    sel = "(fp64 and (fused_moments_vs_oracle or fused_pipeline_against_reference_goldens)) or bit_reproducible or full_size_properties"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-k", sel,
                        "-p", "no:cacheprovider"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
