"""forward_dynamics sanity at several batch sizes: NaN / flag counts, agreement with the oracle on a few systems and between
two identical calls.    python profiles/tools/fd_probe.py [model ...]"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
from bionc_oracle import OracleModel  # noqa: E402
from bionc_b200.bionc_cuda import BiomechanicalModel  # noqa: E402
from bionc_b200.utils import states as st  # noqa: E402

for name in sys.argv[1:] or ["n_link_4", "knee_feikes", "full_body", "lower_limb"]:
    path = os.path.join(REPO, "tests", "golden", "models", name + ".npz")
    model, om = BiomechanicalModel.load(path), OracleModel(path)
    with np.load(os.path.join(REPO, "tests", "golden", "cases", name + ".npz")) as z:
        q0 = np.array(z["Q"][0])
    for B in (16, 33, 1000, 150_001):
        if name == "n_link_4":
            Qn, Qdn = st.planar_chain_states(model.table, min(B, 1 << 14), seed=31)
        elif name == "knee_feikes":
            Qn, Qdn = st.rigidly_moved_states(q0, min(B, 1 << 14), seed=31, angle_max=0.3, omega_sigma=0.5)
        else:
            Qn, Qdn = st.tree_states(model.table, min(B, 1 << 14), seed=31)
        reps = (B + Qn.shape[0] - 1) // Qn.shape[0]
        Q = torch.from_numpy(Qn).cuda().repeat(reps, 1)[:B].contiguous()
        Qd = torch.from_numpy(Qdn).cuda().repeat(reps, 1)[:B].contiguous()
        Qc, Qdc = Q.clone(), Qd.clone()
        a = model.forward_dynamics(Q, Qd, return_status=True, pivoting="never")
        b = model.forward_dynamics(Q, Qd, return_status=True, pivoting="never")
        errs = []
        for s in (0, B // 2, B - 1):
            rq, rl = om.forward_dynamics(Qn[s % Qn.shape[0]], Qdn[s % Qn.shape[0]])
            errs.append(max(float(np.abs(a[0][s].cpu().numpy() - rq).max()), float(np.abs(a[1][s].cpu().numpy() - rl).max())))
        print(f"{name:12s} B={B:6d} layout spc={model.kernel_layout()['systems_per_cta']:2d}  nan qdd {int(torch.isnan(a[0]).any(1).sum()):6d} "
              f"nan lam {int(torch.isnan(a[1]).any(1).sum()):6d} flagged {int((a[2] != 0).sum()):6d}  oracle err {max(errs):.1e}  "
              f"repeat equal {torch.equal(torch.nan_to_num(a[0]), torch.nan_to_num(b[0])) and torch.equal(torch.nan_to_num(a[1]), torch.nan_to_num(b[1]))}  "
              f"inputs untouched {torch.equal(Q, Qc) and torch.equal(Qd, Qdc)}")
