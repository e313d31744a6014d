#!/usr/bin/env python
"""Run GPU test files repeatedly inside ONE process (rare, timing-dependent failures: a long-lived process that has
already run other kernels is the context they were seen in).  usage: repeat_gpu_tests.py REPS file [file ...]"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
reps, files = int(sys.argv[1]), sys.argv[2:]
bad = 0
for rep in range(reps):
    rc = pytest.main(["-q", "-m", "gpu", "-p", "no:cacheprovider", "--no-header", "--tb=short"] + files)
    bad += int(rc != 0)
    print(f"rep {rep}: rc {int(rc)}", flush=True)
print(f"failing repetitions: {bad} of {reps}")
sys.exit(1 if bad else 0)
