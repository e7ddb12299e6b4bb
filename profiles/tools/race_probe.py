"""Small repeatability / race probe: which outputs of repeated identical calls differ, for which systems.
    python profiles/tools/race_probe.py [model] [batch]          (run it under compute-sanitizer --tool racecheck with a small batch)"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from bionc_b200.bionc_cuda import BiomechanicalModel  # noqa: E402
from bionc_b200.utils import states as st  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "n_link_4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 150_001
model = BiomechanicalModel.load(os.path.join(REPO, "tests", "golden", "models", name + ".npz"))
if name == "n_link_4":
    Qn, Qdn = st.planar_chain_states(model.table, min(B, 1 << 14), seed=31)
else:
    Qn, Qdn = st.tree_states(model.table, min(B, 1 << 14), seed=31)
reps = (B + Qn.shape[0] - 1) // Qn.shape[0]
Q = torch.from_numpy(Qn).cuda().repeat(reps, 1)[:B].contiguous()
Qd = torch.from_numpy(Qdn).cuda().repeat(reps, 1)[:B].contiguous()
print(model.kernel_layout())


def diff(tag, a, b):
    for k, (x, y) in enumerate(zip(a, b)):
        if not torch.equal(x, y):
            bad = (x != y).reshape(x.shape[0], -1).any(dim=1).nonzero().flatten()
            d = (x - y).abs().max().item()
            print(f"{tag}: output {k} differs in {bad.numel()} systems, max |d| {d:.3e}, first {bad[:12].tolist()}, mod 32 {sorted(set((bad % 32).tolist()))[:32]}")
            return
    print(f"{tag}: identical")


for tag, fn in (
    ("forward_dynamics", lambda: model.forward_dynamics(Q, Qd, pivoting="never")),
    ("rk4 20 steps", lambda: model.rk4(Q, Qd, h=1e-3, n_steps=20)),
    ("rk4 1 step", lambda: model.rk4(Q, Qd, h=1e-3, n_steps=1)),
    ("rk4 20 steps + lambdas", lambda: model.rk4(Q, Qd, h=1e-3, n_steps=20, return_lambdas=True)),
):
    ref = fn()
    for _ in range(3):
        diff(tag, ref, fn())
