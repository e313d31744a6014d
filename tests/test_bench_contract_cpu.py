"""bench.py contract on CPU: the reference arm (oracle port on host cores) prints ONE JSON line with the keys the
driver reads, and the same `config` object as the B200 arm builds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-ncoarse", "5"], capture_output=True, text=True, timeout=600)
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
    assert d["config"] == bench.workload_config(9)      # what the B200 arm reports for the default workload


def test_every_timed_kernel_has_an_algorithmic_byte_entry_or_is_overhead():
    sys.path.insert(0, ROOT)
    import bench
    step_kernels = ["psc_kick_drift_wrap_count", "psc_bin_particles_counted", "psc_deposit_binned", "psc_fft_r2c",
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
