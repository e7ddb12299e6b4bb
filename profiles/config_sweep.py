"""Device-resident throughput of the fused integrator for every configuration of BASELINE.json
(bench.py measures the headline one, config 3).  Writes one JSON object per configuration.

    python profiles/config_sweep.py [--batch 1000000] [--out gpurun_out/config_sweep.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bionc_b200.bionc_cuda import BiomechanicalModel, ExternalForceSet, _lib  # noqa: E402
from bionc_b200.utils import roofline as rf  # noqa: E402
from bionc_b200.utils import states as st  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def case(name):
    with np.load(os.path.join(GOLD, "cases", name + ".npz")) as z:
        return {k: np.array(z[k]) for k in z.files}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--steps-per-launch", type=int, default=10)
    ap.add_argument("--launches", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "config_sweep.json"))
    ap.add_argument("--no-specialize", action="store_true", help="general-path models: the generic kernel")
    ap.add_argument("--only", default=None, help="substring of the model name to restrict the sweep to")
    args = ap.parse_args()
    B, S = args.batch, args.steps_per_launch
    lib = _lib.load()
    best, sus = C.c_double(), C.c_double()
    _lib.check(lib.bionc_cuda_measure_fp64_peak(0, 10, C.byref(best), C.byref(sus)), "peak")
    results = []
    configs = [
        ("config1 pendulum.nmod (ground hinge, indefinite G)", "pendulum_root", "root", 0.01, None, False),
        ("config2 pendulum_with_force.nmod + no_equilibrium force", "pendulum_force_batch", "hinge", 1.0 / 60.0, "case", False),
        ("config3 lower limb (4 segments, 3 spherical)", "lower_limb", "tree", 0.01, None, False),
        ("config4 knee Feikes (closed loops)", "knee_feikes", "knee", 5e-5, None, False),
        ("4-link hinge pendulum", "n_link_4", "chain", 0.005, None, False),
        ("config5 full body (10 segments), shared inertia", "full_body", "tree", 0.01, None, False),
        ("config5 full body (10 segments), per-system inertial parameters", "full_body", "tree", 0.01, None, True),
    ]
    for label, name, kind, h, fx, per_sys in configs:
        if args.only is not None and args.only not in name:
            continue
        model = BiomechanicalModel.load(os.path.join(GOLD, "models", name + ".npz"))
        cs = case(name)
        pool = min(B, 1 << 15)
        if kind == "tree":
            Q, Qd = st.tree_states(model.table, pool, seed=1)
        elif kind == "hinge":
            Q, Qd = st.hinge_pendulum_states(pool, seed=1)
        elif kind == "chain":
            Q, Qd = st.planar_chain_states(model.table, pool, seed=1)
        elif kind == "knee":
            Q, Qd = st.rigidly_moved_states(cs["Q"][0], pool, seed=1, angle_max=0.3, omega_sigma=0.0)
        else:
            Q, Qd = np.repeat(cs["Q"][:1], pool, 0), np.repeat(cs["Qdot"][:1], pool, 0)
        reps = (B + pool - 1) // pool
        Qt = torch.from_numpy(Q).cuda().repeat(reps, 1)[:B].contiguous()
        Qdt = torch.from_numpy(Qd).cuda().repeat(reps, 1)[:B].contiguous()
        fext = None
        if fx == "case":
            fext = ExternalForceSet.empty_from_nb_segment(model.nb_segments)
            for s, k, p in zip(cs["fext_seg"], cs["fext_kind"], cs["fext_param"]):
                fext.external_forces[int(s)].append((int(k), np.array(p)))
        ip = None
        if per_sys:
            mass, com, J = st.perturbed_inertial_parameters(model.table, pool, seed=3)
            ipn = np.concatenate([mass[..., None], com, np.stack([J[..., 0, 0], J[..., 0, 1], J[..., 0, 2], J[..., 1, 1], J[..., 1, 2], J[..., 2, 2]], -1)], -1)
            ip = torch.from_numpy(ipn).cuda().repeat(reps, 1, 1)[:B].contiguous()
        # general-path models run their own build of the kernel, like bench.py's workloads
        general = fext is not None or any(int(t) != 0 for t in model.table.blk_type)
        spec_s = model.specialize(external_forces=fext, per_system_inertia=per_sys) if (general and not args.no_specialize) else None
        kw = dict(h=h, n_steps=S, normalize=True, external_forces=fext, inertial_parameters=ip, inplace=False, specialize=False)
        for _ in range(3):
            model.rk4(Qt, Qdt, **kw)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.launches)]
        for e0, e1 in ev:
            Qw, Qdw = Qt.clone(), Qdt.clone()
            opt, keep = model._options(fext, None, ip, B)
            status = torch.zeros(B, dtype=torch.int32, device="cuda")
            e0.record()
            rc = lib.bionc_cuda_rk4(model._h, Qw.data_ptr(), Qdw.data_ptr(), h, None, S, 1, C.byref(opt), None, 1, None,
                                    status.data_ptr(), B, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            e1.record()
            _lib.check(rc, "rk4")
        torch.cuda.synchronize()
        ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
        rate = B * S / (ms * 1e-3)
        fl = rf.flops_per_step(model.table)
        res = {
            "config": label, "model": name, "segments": model.nb_segments, "m": model.nb_holonomic_constraints,
            "systems": B, "rk4_steps_per_launch": S, "h": h, "kernel_ms": ms, "system_steps_per_s": rate,
            "flops_per_system_step": fl, "algorithmic_tflops": rate * fl / 1e12,
            "frac_of_measured_fp64_peak": rate * fl / 1e12 / sus.value, "fp64_peak_tflops": sus.value,
            "flagged": int((status != 0).sum().item()), "layout": model.kernel_layout(),
            "kernel_build": "generic" if spec_s is None else "model-specialised (NVRTC, %.1f s)" % spec_s,
        }
        print(json.dumps(res), flush=True)
        results.append(res)
        del model, Qt, Qdt
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
