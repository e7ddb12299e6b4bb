"""Achieved HBM bandwidth of the component / post-processing evaluators (bionc_cuda_eval) on the lower limb,
1e6 systems resident on the device: algorithmic bytes (input read once + output written once) / CUDA-event time.
Usage (GPU box): python profiles/post_sweep.py > gpurun_out/post_sweep.json"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)
from bionc_b200.bionc_cuda import BiomechanicalModel  # noqa: E402
from bionc_b200.utils import states as st  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    times = []
    for _ in range(reps):
        flush.zero_()  # evict L2 between repetitions
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.median(times))  # the caching allocator occasionally calls cudaMalloc inside a repetition


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    model = BiomechanicalModel.load(os.path.join(REPO, "tests", "golden", "models", "lower_limb.npz"))
    Q, Qd = st.tree_states(model.table, B, seed=0)
    Qt, Qdt = torch.from_numpy(Q).cuda(), torch.from_numpy(Qd).cuda()
    n, m, N, nmk = model.nb_Q, model.nb_holonomic_constraints, model.nb_segments, model.nb_markers
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6550.0))
    rows = []
    cases = [
        ("markers (forward kinematics)", lambda: model.markers(Qt), n + 3 * nmk),
        ("center_of_mass_position", lambda: model.center_of_mass_position(Qt), n + 3 * N),
        ("kinetic_energy", lambda: model.kinetic_energy(Qdt), n + 1),
        ("potential_energy", lambda: model.potential_energy(Qt), n + 1),
        ("constraint_residual_norms", lambda: model.constraint_residual_norms(Qt), n + 2),
        ("holonomic_constraints", lambda: model.holonomic_constraints(Qt), n + m),
        ("holonomic_constraints_acceleration_bias", lambda: model.holonomic_constraints_acceleration_bias(Qdt), n + m),
    ]
    mj = model.nb_joint_constraints
    cases += [
        ("rigid_body_constraints", lambda: model.rigid_body_constraints(Qt), n + 6 * N),
        ("joint_constraints", lambda: model.joint_constraints(Qt), n + mj),
        # the single-evaluation kernel behind the Python-level RK4 closure: Q, Qdot in; Qddot, lambda, status out
        ("forward_dynamics (one evaluation, pivoting='never')", lambda: model.forward_dynamics(Qt, Qdt, pivoting="never"), 3 * n + m + 0.5),
    ]
    only = os.environ.get("POST_SWEEP_ONLY")
    if only:
        cases = [c for c in cases if only in c[0]]
    for name, fn, doubles in cases:
        ms = timed(fn)
        gbs = B * doubles * 8 / (ms * 1e-3) / 1e9
        rows.append({"evaluator": name, "systems": B, "ms": ms, "algorithmic_bytes": B * doubles * 8, "GB_per_s": gbs, "frac_of_hbm_peak": gbs / hbm, "hbm_peak_GB_per_s": hbm})
        print(json.dumps(rows[-1]))
    # the Jacobians are dense (B, rows, 12N) outputs: 1e5 systems are 1.3 GB for K_h, ten times the L2
    if only:
        return
    Bk = min(B, 100_000)
    Qk = Qt[:Bk].contiguous()
    for name, fn, doubles in [
        ("rigid_body_constraints_jacobian", lambda: model.rigid_body_constraints_jacobian(Qk), n + 6 * N * n),
        ("joint_constraints_jacobian", lambda: model.joint_constraints_jacobian(Qk), n + mj * n),
        ("holonomic_constraints_jacobian", lambda: model.holonomic_constraints_jacobian(Qk), n + m * n),
    ]:
        ms = timed(fn)
        gbs = Bk * doubles * 8 / (ms * 1e-3) / 1e9
        rows.append({"evaluator": name, "systems": Bk, "ms": ms, "algorithmic_bytes": Bk * doubles * 8, "GB_per_s": gbs, "frac_of_hbm_peak": gbs / hbm, "hbm_peak_GB_per_s": hbm})
        print(json.dumps(rows[-1]))


if __name__ == "__main__":
    main()
