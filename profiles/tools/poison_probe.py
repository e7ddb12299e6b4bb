"""Which region of the shared-memory working set is read before it is written?  Every golden model, forward_dynamics and
a 3-step rk4, with each region filled with NaN in turn (bionc_cuda_debug_poison)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from bionc_b200.bionc_cuda import BiomechanicalModel, _lib  # noqa: E402

lib = _lib.load()
names = sorted(f[:-4] for f in os.listdir(os.path.join(REPO, "tests", "golden", "models")) if f.endswith(".npz"))
REG = ["S0", "S1", "r0", "r1", "lam", "dinv", "knl"]
for name in names:
    model = BiomechanicalModel.load(os.path.join(REPO, "tests", "golden", "models", name + ".npz"))
    with np.load(os.path.join(REPO, "tests", "golden", "cases", name + ".npz")) as z:
        Q, Qd = np.array(z["Q"]), np.array(z["Qdot"])
    Q, Qd = np.tile(Q, (40, 1))[:67], np.tile(Qd, (40, 1))[:67]
    lib.bionc_cuda_debug_poison(0)
    ref = model.forward_dynamics(Q, Qd, pivoting="never") + model.rk4(Q, Qd, h=1e-4, n_steps=3) + \
        model.forward_dynamics(Q, Qd, pivoting="never", stabilization=dict(alpha=1.0, beta=1.0))
    bad = []
    for bit in range(7):
        lib.bionc_cuda_debug_poison(1 << bit)
        out = model.forward_dynamics(Q, Qd, pivoting="never") + model.rk4(Q, Qd, h=1e-4, n_steps=3) + \
            model.forward_dynamics(Q, Qd, pivoting="never", stabilization=dict(alpha=1.0, beta=1.0))
        which = [k for k, (a, b) in enumerate(zip(ref, out)) if not np.array_equal(a, b, equal_nan=True)]
        if which:
            bad.append(f"{REG[bit]}:{which}")
    lib.bionc_cuda_debug_poison(0)
    print(f"{name:44s} {model.kernel_layout()['serial_joint_level_schedule']} {'clean' if not bad else ' '.join(bad)}")
