#!/usr/bin/env python
"""The reference's 1_test_solve.py against this repo's build: XM.solve on the shipped SIMPLE1 matrix (447 x 447).
Run after `python __graft_entry__.py` on a B200:  python examples/1_solve_simple1.py [workdir]"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.append(os.path.join(ROOT, "XM", "build"))
sys.path.insert(0, ROOT)

import XM  # noqa: E402  (the compiled pybind11 module, same surface as the reference's)

from xm_code_b200.binio import load_matrix_from_bin  # noqa: E402

work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/xm_simple1"
os.makedirs(work, exist_ok=True)
shutil.copyfile(os.path.join(ROOT, "tests", "golden", "simple1_Q.bin"), os.path.join(work, "Q.bin"))

# full XM: rank staircase from 3, certificate after every rank (reference call: 1_test_solve.py:42)
XM.solve(work + "/", 3, 1e-16, 0.0, 1000)

R = load_matrix_from_bin(work + "/R.bin")
s = load_matrix_from_bin(work + "/s.bin")
print("R", R.shape, "s", s.shape, "s range", float(s.min()), float(s.max()))
