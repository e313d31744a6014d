"""bench.py contract on CPU: the reference arm (oracle port on host cores) prints ONE JSON line with the keys the
driver reads, and the same `config` object as the B200 arm builds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ncoarse", "5", "--cpu-ncoarse", "5"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_updates_per_sec_full_pm_step"
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(5)      # the config names the size that actually ran
    assert d["scaling"] == "strong"


def test_reference_arm_names_the_sample_it_ran():
    """a workload whose K + W steps do not fit the time budget is sampled at --cpu-ncoarse, and the line says so"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ncoarse", "6", "--cpu-ncoarse", "5", "--cpu-budget-s", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["config"]["ncells_1d"] == 32 and "32^3" in d["config"]["workload"]
    assert "1/8 of the 64^3 workload" in d["cpu_baseline"]["sample"]


def test_ics_are_the_same_in_both_arms():
    """bench.lattice_ics: the NumPy (CPU arm) and the torch (GPU arms, any x-slab) flavours give the same particles"""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    N = 16
    p0, v0, i0 = bench.lattice_ics(np, N, 0, N)
    p1, v1, i1 = bench.lattice_ics(torch, N, 4, 8, device="cpu")
    sl = slice(4 * N * N, 12 * N * N)
    assert np.array_equal(i0[sl], i1.numpy())
    assert np.abs(p0[sl] - p1.numpy()).max() < 2e-7 and np.abs(v0[sl] - v1.numpy()).max() < 1e-9
    assert 0.0 <= p0.min() and p0.max() < 1.0
    assert abs(v0.std() / 1e-3 - 1) < 0.02 and abs(v0.mean()) < 1e-4
    d = (p0 * N - (np.stack(np.meshgrid(*[np.arange(N) + 0.5] * 3, indexing="ij"), -1).reshape(-1, 3)))
    d = d - N * np.round(d / N)
    assert abs(d.std() / 0.3 - 1) < 0.03


def test_every_timed_kernel_has_an_algorithmic_byte_entry_or_is_overhead():
    sys.path.insert(0, ROOT)
    import bench
    step_kernels = ["psc_step_sort", "psc_deposit_sorted", "psc_interp_kick_phi_sorted",
                    "psc_kick_drift_wrap_count", "psc_bin_particles_counted", "psc_deposit_binned", "psc_fft_r2c",
                    "psc_green", "psc_fft_c2r", "psc_interp_kick_phi_binned", "psc_kick_drift_wrap_slab",
                    "psc_bin_particles_slab", "psc_deposit_binned_slab", "psc_interp_kick_phi_binned_slab",
                    "psc_slab_fft_r2c_planes", "psc_slab_fft_x", "psc_green_slab", "psc_slab_fft_c2r_planes"]
    missing = [k for k in step_kernels if k not in bench.ALGO_BYTES]
    assert not missing, missing
    # the per-step algorithmic budget of SURVEY 8(d)
    assert bench.STEP_ALGO_BYTES == 176.0
    single = ["psc_kick_drift_wrap_count", "psc_deposit_binned", "psc_fft_r2c", "psc_green", "psc_fft_c2r",
              "psc_interp_kick_phi_binned"]
    assert sum(bench.ALGO_BYTES[k] for k in single) == 176.0
    loop = ["psc_step_sort", "psc_deposit_sorted", "psc_fft_r2c", "psc_green", "psc_fft_c2r", "psc_interp_kick_phi_sorted"]
    assert sum(bench.ALGO_BYTES[k] for k in loop) == 176.0
    fused = ["psc_step_sort", "psc_deposit_sorted", "psc_fft_poisson", "psc_interp_kick_phi_sorted"]
    assert sum(bench.ALGO_BYTES[k] for k in fused) == 176.0
