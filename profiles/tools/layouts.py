"""Kernel layout of the golden models (diagnostics): python profiles/tools/layouts.py"""
import glob
import os
import sys

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, REPO)
from bionc_b200.bionc_cuda import BiomechanicalModel  # noqa: E402

for p in sorted(glob.glob(os.path.join(REPO, "tests", "golden", "models", "*.npz"))):
    name = os.path.basename(p)[:-4]
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    m = BiomechanicalModel.load(p)
    Q = np.zeros((1, m.nb_Q))
    try:
        m.forward_dynamics(Q + 0.1 * np.arange(m.nb_Q), Q, pivoting="never")
    except Exception:
        pass
    L = m.kernel_layout()
    print(name, L)
