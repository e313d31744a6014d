"""The float32 fast path of the f(R) cubic root (csrc/fr_roots.cuh: psc::solve_cubic) against the float64 statement of
cubic.py:162-207 (psc::solve_cubic_f64), on the host build of the very same header: 3 M log-uniform (p, d1) pairs of
all sign combinations and 3 M pairs of the f(R) regime.  The two must agree to float32 rounding (3e-7) and must have
their NaNs (the reference's pow(negative, 1/3)) in the same places."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cubic_fast_path_matches_float64_statement(tmp_path):
    exe = str(tmp_path / "fr_roots_harness")
    subprocess.check_call(["g++", "-O2", "-I", os.path.join(ROOT, "pysco_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "fr_roots_harness.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, check=True).stdout
    rows = re.findall(r"worst rel diff ([0-9.e+-]+), >2e-6: (\d+), NaN mismatches (\d+)", out)
    assert len(rows) == 2, out
    for worst, big, nan in rows:
        assert float(worst) < 3e-7 and int(big) == 0 and int(nan) == 0, out
