#!/usr/bin/env python
"""bench.py -- RK4 forward-dynamics throughput of the bionc_cuda backend (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config 3 of BASELINE.json): the 4-segment lower-limb model (PELVIS free, hip / knee /
ankle spherical; geometry of the reference's examples/models/lower_limb.nc, inertia of SURVEY.md
8(d)), B = 1e6 independent systems PER GPU with seeded, constraint-consistent, pairwise distinct random
initial states, h = 0.01 s, post-step renormalisation on.  One bench "step" = one launch of the fused
integrator advancing every system by `rk4_steps_per_launch` (default 100) RK4 steps, 4 forward-dynamics
evaluations each; the states are put back to the initial ones every 1000 steps, so the timed region is a
sequence of the configuration's own 1000-step integrations (10 launches each).
Metric: RK4 system-steps/s, whole job (all GPUs).

  value        states resident in HBM, CUDA events on the launching stream, max over ranks
  e2e          the same through the C ABI's host-buffer entry point (bionc_cuda_rk4_host): pinned
               host arrays in, H2D + kernel + D2H inside the timed region
  roofline     bound: the FP64 pipe.  `frac` = FP64 operations the kernel EXECUTES (ncu opcode counts of a committed
               capture of this kernel and configuration: 2 DFMA + DMUL + DADD per system-step) / measured kernel time /
               the DFMA peak measured in this run.  `algorithmic_frac` is the same with SURVEY.md 8(d)'s operation count
               of the reference-shaped algorithm (dense m^3/3 factorisation), which the kernel does not execute: it can
               exceed 1.  `traffic` (ncu DRAM bytes per launch) against the algorithmic bytes of a launch (the state
               read once and written once: it stays on chip for all steps of a launch) gives `traffic_ratio`.
  cpu_baseline the UNMODIFIED reference (bionc.RK4 + bionc_numpy forward_dynamics, installed in baseline/_ref) on
               this box's host cores, one process per core (N = 1, rank 0); falls back to the numpy oracle port
               (kind "port") when the reference install is missing
  stabilized   the same launch with Baumgarte stabilisation (what long simulations use): rate and flagged systems
  strong       (N > 1) the configuration's own scaling run: B = 1e6 systems IN TOTAL, 1e6 / N per GPU
  gather       (N > 1) the one collective of the workload: all_gather_into_tensor of the final states

`--impl reference` times the reference's own CPU implementation on all host cores (bounded sample per step).
"""

from __future__ import annotations

import argparse
import json
import os

# the CPU legs run one worker process per core; BLAS threading is useless at 81 x 81 and thrashes
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOAD = "lower_limb_4seg (BASELINE.json configs[2]; PELVIS free + 3 spherical joints, n=48, m=33)"
# The default workload is the one BASELINE.json's metric is quoted on (the 4-segment lower limb).  `--workload`
# runs the other BASELINE configurations through the same measurement (SURVEY 8(d) definitions); they are not
# bench lines of the driver, which calls bench.py without that flag.  `ref`: the configuration's builder in
# baseline/reference_models.py (the live reference).
WORKLOADS = {
    "lower_limb": dict(model="lower_limb", ref="lower_limb", states="tree", h=0.01, fext=False, inertia=False, desc=WORKLOAD),
    "pendulum_force": dict(model="pendulum_force_batch", ref="pendulum_force", states="hinge", h=1.0 / 60.0, fext=True, inertia=False, specialize=True,
                           desc="pendulum_with_force.nmod + no_equilibrium force (BASELINE.json configs[1]; ground hinge, n=12, m=11)"),
    "knee": dict(model="knee_feikes", ref="knee", states="knee", h=5e-5, fext=False, inertia=False, specialize=True,
                 desc="knee_feikes parallel mechanism (BASELINE.json configs[3]; closed loops, n=24, m=20)"),
    "full_body": dict(model="full_body", ref="full_body", states="tree", h=0.001, fext=False, inertia=True,
                      desc="synthetic full body, per-system inertial parameters (BASELINE.json configs[4]; 10 segments, n=120, m=87)"),
    "n_link_4": dict(model="n_link_4", ref=None, states="chain", h=0.005, fext=False, inertia=False, specialize=True,
                     desc="4-link hinge pendulum (tests/test_forward_dynamics.py of the reference; 4 hinges, n=48, m=44)"),
}
STAB = dict(alpha=50.0, beta=20.0)  # Baumgarte gains of the stabilised leg: examples/forward_dynamics/actuated_3d_pendulum.py:155

# What one launch of the dominant kernel does on the device, from the committed `ncu --set full` captures of THIS kernel
# version at the bench configuration (B = 1e6 systems, 100 steps per launch): executed FP64 thread instructions
# (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum), all warp instructions, FP64 pipe utilisation and
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum).  Filled from profiles/ by profiles/ncu_raw_summary.py.
NCU = {}
try:
    with open(os.path.join(REPO, "profiles", "ncu_constants.json")) as _f:
        NCU = json.load(_f)
except OSError:
    pass

METRIC = "rk4_forward_dynamics_system_steps_per_s"
UNIT = "system-steps/s"


def workload_paths(name):
    w = WORKLOADS[name]
    return (os.path.join(REPO, "tests", "golden", "models", w["model"] + ".npz"),
            os.path.join(REPO, "tests", "golden", "cases", w["model"] + ".npz"))


def workload_states(name, table, count, seed):
    """Seeded consistent initial states of a workload (numpy, SURVEY 8(d) generators); every system its own draw."""
    from bionc_b200.utils import states as st

    kind = WORKLOADS[name]["states"]
    if kind == "tree":
        return st.tree_states(table, count, seed=seed)
    if kind == "hinge":
        return st.hinge_pendulum_states(count, seed=seed)
    if kind == "chain":
        return st.planar_chain_states(table, count, seed=seed)
    with np.load(workload_paths(name)[1]) as z:  # knee: rigid motions of the reference's own initial state
        q0 = np.array(z["Q"][0])
    return st.rigidly_moved_states(q0, count, seed=seed, angle_max=0.3, omega_sigma=0.0)


def workload_fext(name):
    """(segment, kind, param) arrays of the workload's external forces (from the golden case), or None."""
    if not WORKLOADS[name]["fext"]:
        return None
    with np.load(workload_paths(name)[1]) as z:
        return np.array(z["fext_seg"]), np.array(z["fext_kind"]), np.array(z["fext_param"])


# ----------------------------------------------------------------------------- CPU side: the reference on the host cores
def reference_available(workload) -> bool:
    sys.path.insert(0, os.path.join(REPO, "baseline"))
    import reference_models as rm

    return rm.available() and WORKLOADS[workload]["ref"] is not None


def _cpu_worker(args):
    """One host core: integrate `nsys` systems for `nsteps` RK4 steps.  kind "reference": the unmodified reference
    (bionc.RK4 around model.forward_dynamics, utils/ode_solver.py:8-53, :107-116); kind "port": the numpy oracle."""
    seed, nsys, nsteps, workload, kind = args
    os.environ["OMP_NUM_THREADS"] = "1"
    warnings.filterwarnings("ignore")
    from bionc_b200.bionc_cuda.model_table import ModelTable
    from bionc_b200.utils import states as st

    w = WORKLOADS[workload]
    model_file = workload_paths(workload)[0]
    table = ModelTable.load(model_file)
    Q, Qd = workload_states(workload, table, nsys, seed)
    t = np.arange(nsteps + 1) * w["h"]
    per_system = st.perturbed_inertial_parameters(table, nsys, seed=seed) if w["inertia"] else None
    if kind == "reference":
        sys.path.insert(0, os.path.join(REPO, "baseline"))
        import reference_models as rm

        model, fext = rm.CONFIGS[w["ref"]]["build"]()
        t0 = time.perf_counter()
        for s in range(nsys):
            if per_system is not None:  # the reference carries the inertial parameters in the model object (inside the timing)
                for i, seg in enumerate(model.segments_no_ground.values()):
                    seg.set_natural_inertial_parameters(mass=per_system[0][s, i], natural_center_of_mass=per_system[1][s, i],
                                                        natural_pseudo_inertia=per_system[2][s, i])
                model._update_mass_matrix()
            rm.reference_rk4(model, Q[s], Qd[s], t, True, fext)
        return time.perf_counter() - t0
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from bionc_oracle import OracleModel  # noqa: E402  (the CPU baseline leg is allowed to run the oracle)

    om = OracleModel(model_file)
    fx = workload_fext(workload)
    fext = None if fx is None else [(int(a), int(b), c) for a, b, c in zip(*fx)]
    t0 = time.perf_counter()
    for s in range(nsys):
        if per_system is not None:
            om = OracleModel(model_file, seg_mass=per_system[0][s], seg_com=per_system[1][s], seg_J=per_system[2][s])
        om.rk4(Q[s], Qd[s], t, normalize=True, fext=fext)
    return time.perf_counter() - t0


def cpu_rate(nsys_per_core: int, nsteps: int, cores: int, seed0: int = 1000, workload: str = "lower_limb", kind: str = "reference"):
    """Aggregate system-steps/s of `cores` worker processes, wall clock of the slowest worker."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(seed0 + c, nsys_per_core, nsteps, workload, kind) for c in range(cores)])
    wall = max(times)  # integration time of the slowest worker (process start-up, imports and model building excluded)
    return cores * nsys_per_core * nsteps / wall, wall


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_kind(workload):
    return "reference" if reference_available(workload) else "port"


CPU_WHAT = {"reference": "unmodified bioNC: bionc.RK4 + bionc_numpy.BiomechanicalModel.forward_dynamics (baseline/_ref)",
            "port": "numpy oracle port of the reference algorithm (dense KKT, LAPACK LU); baseline/_ref missing"}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on all host cores.  Each bench step is a bounded
    sample: every core integrates `nsys` systems of the workload for `nsteps` RK4 steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    kind = cpu_kind(args.workload)
    nsys, nsteps = (2, 25) if kind == "reference" else (4, 100)  # ~0.4 s per core per bench step
    if WORKLOADS[args.workload]["model"] == "full_body":
        nsys, nsteps = 1, 20
    for _ in range(args.warmup):
        cpu_rate(nsys, nsteps, cores, workload=args.workload, kind=kind)
    wall = 0.0
    for k in range(args.steps):
        wall += cpu_rate(nsys, nsteps, cores, seed0=2000 + 100 * k, workload=args.workload, kind=kind)[1]
    value = args.steps * cores * nsys * nsteps / wall
    sample = f"{cores} cores x {nsys} systems x {nsteps} RK4 steps per bench step ({CPU_WHAT[kind]})"
    wl = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "systems_per_gpu": args.batch, "rk4_steps_per_launch": args.rk4_steps_per_launch,
                   "h": wl["h"], "normalize": True,
                   "sample": "the CPU arm integrates a subsample of the same workload (same model, h, normalize, state generator)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU side
class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  The sampler is started before the
    warm-up (nvidia-smi takes a moment to come up, longer with 8 GPUs busy) and every row is stamped on
    arrival; `summary(t0, t1)` keeps the rows that fall inside the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        sm, mx, reasons = [], 0.0, set()
        for ts, r in self.rows:
            if not (t0 <= ts <= t1 + 0.05):
                continue
            try:
                sm.append(float(r[0])), (mx := max(mx, float(r[1])))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1

    # the CPU baseline runs first, before CUDA is initialised in this process (fork safety)
    cpu_baseline = None
    if not distributed and not args.no_cpu_baseline:
        cores = host_cores()
        kind = cpu_kind(args.workload)
        nsys, nsteps = (4, 100) if kind == "reference" else (8, 1000)  # ~3-25 s of CPU work per core
        if WORKLOADS[args.workload]["model"] == "full_body":
            nsys, nsteps = (2, 50) if kind == "reference" else (4, 250)
        rate, wall = cpu_rate(nsys, nsteps, cores, workload=args.workload, kind=kind)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{cores} cores x {nsys} systems x {nsteps} RK4 steps, same model/h/normalize/state generator "
                                  f"({wall:.1f} s wall; {CPU_WHAT[kind]})"}

    import ctypes as C

    import torch

    from bionc_b200.bionc_cuda import BiomechanicalModel, _lib
    from bionc_b200.utils import roofline as rf
    from bionc_b200.utils import states as st

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    B, S = args.batch, args.rk4_steps_per_launch
    wl = WORKLOADS[args.workload]
    h_step = wl["h"]
    model = BiomechanicalModel.load(workload_paths(args.workload)[0], device=dev)
    n = model.nb_Q
    # seeded synthetic initial states: B distinct consistent states per rank (numpy generators of SURVEY 8(d))
    Qp, Qdp = workload_states(args.workload, model.table, B, seed=1 + rank)
    Q0 = torch.from_numpy(Qp).to(dev)
    Qd0 = torch.from_numpy(Qdp).to(dev)
    Q, Qd = Q0.clone(), Qd0.clone()
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    lib = _lib.load()
    opt = _lib.BioncDynOptions()
    keep = []
    fx = workload_fext(args.workload)
    if fx is not None:
        seg, kind, par = (np.ascontiguousarray(fx[0], dtype=np.int32), np.ascontiguousarray(fx[1], dtype=np.int32),
                          np.ascontiguousarray(fx[2], dtype=np.float64))
        fdesc = _lib.BioncFextDesc(int(seg.size), seg.ctypes.data_as(_lib.c_int32_p), kind.ctypes.data_as(_lib.c_int32_p),
                                   par.ctypes.data_as(_lib.c_double_p))
        keep += [seg, kind, par, fdesc]
        opt.fext = C.pointer(fdesc)
    ip_host = None
    if wl["inertia"]:  # config 5: per-system natural inertial parameters (B, N, 10), resident on the device
        mass, com, J = st.perturbed_inertial_parameters(model.table, B, seed=100 + rank)
        ip_host = np.ascontiguousarray(np.concatenate(
            [mass[..., None], com, np.stack([J[..., 0, 0], J[..., 0, 1], J[..., 0, 2], J[..., 1, 1], J[..., 1, 2], J[..., 2, 2]], axis=-1)], axis=-1))
        ip_dev = torch.from_numpy(ip_host).to(dev)
        keep.append(ip_dev)
        opt.inertia = ip_dev.data_ptr()
    stream = torch.cuda.current_stream(dev)
    # General-path workloads run the model's own build of the integrator (bionc_cuda_model_specialize: the model tables as
    # compile-time constants, NVRTC): a per-model compilation step like model_create, outside the timed region.  The lean
    # workloads (lower limb, full body) keep the generic kernel, which is faster for them.
    spec_seconds = None
    if wl.get("specialize") and not args.no_specialize:
        t_spec = time.perf_counter()
        _lib.check(lib.bionc_cuda_model_specialize(model._h, C.byref(opt), 0, 1), "model_specialize")
        sopt = _lib.BioncDynOptions()
        C.memmove(C.byref(sopt), C.byref(opt), C.sizeof(opt))
        sopt.use_stabilization = 1  # the stabilised leg's shape (the same kernel on the general path)
        _lib.check(lib.bionc_cuda_model_specialize(model._h, C.byref(sopt), 0, 1), "model_specialize")
        spec_seconds = time.perf_counter() - t_spec

    def launch(o=opt, q=Q, qd=Qd, b=B):
        rc = lib.bionc_cuda_rk4(model._h, q.data_ptr(), qd.data_ptr(), h_step, None, S, 1, C.byref(o), None, 1, None,
                                status.data_ptr(), b, C.c_void_p(stream.cuda_stream))
        _lib.check(rc, "rk4")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(o=opt, q=Q, qd=Qd, b=B, steps=None):
        """`steps` launches, each bracketed by CUDA events on the launching stream; the states go back to the initial
        ones every `run` launches (one integration of the configuration's step count).  Returns (total device ms from the
        first launch to the last, mean ms per launch, flagged systems at the end of the last complete integration)."""
        steps = steps or args.steps
        run = max(1, args.run_steps // S)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        q.copy_(Q0[:b]), qd.copy_(Qd0[:b])
        flagged = torch.zeros((), dtype=torch.int64, device=dev)
        barrier()
        reset_ms = 0.0
        resets = []
        for k, (e0, e1) in enumerate(evs):
            if k > 0 and k % run == 0:  # next integration: back to the initial states (untimed copy, bracketed by its own events)
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record(stream)
                q.copy_(Q0[:b]), qd.copy_(Qd0[:b])
                r1.record(stream)
                resets.append((r0, r1))
            e0.record(stream)
            launch(o, q, qd, b)
            e1.record(stream)
            if (k + 1) % run == 0 or k == steps - 1:
                flagged = (status[:b] != 0).sum()  # flags are per launch: the count after the last launch of an integration
        barrier()
        reset_ms = sum(r0.elapsed_time(r1) for r0, r1 in resets)
        total_ms = evs[0][0].elapsed_time(evs[-1][1]) - reset_ms
        return total_ms, float(np.mean([e0.elapsed_time(e1) for e0, e1 in evs])), int(flagged.item())

    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(args.warmup):
        launch()
    barrier()
    launches0 = lib.bionc_cuda_launch_count()
    t_wall0 = time.perf_counter()
    total_ms, kernel_ms_mean, bad = timed()
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0
    time.sleep(0.1)  # let the last samples of the window arrive
    clocks.__exit__()
    gpu_launches = int(lib.bionc_cuda_launch_count() - launches0)

    # ---- the same launches with Baumgarte stabilisation (biomechanical_model.py:416-421)
    opt_stab = _lib.BioncDynOptions()
    opt_stab.fext, opt_stab.inertia = opt.fext, opt.inertia
    opt_stab.use_stabilization, opt_stab.alpha, opt_stab.beta = 1, STAB["alpha"], STAB["beta"]
    launch(opt_stab)
    stab_steps = max(2, min(args.steps, max(1, args.run_steps // S)))
    stab_total_ms, stab_kernel_ms, stab_bad = timed(opt_stab, steps=stab_steps)

    # ---- end-to-end through the C ABI with host buffers (pinned), copies inside the timed region
    Be = min(B, args.e2e_batch)
    Qh = torch.from_numpy(np.ascontiguousarray(Qp[:Be])).pin_memory()
    Qdh = torch.from_numpy(np.ascontiguousarray(Qdp[:Be])).pin_memory()
    Qh0, Qdh0 = Qh.clone(), Qdh.clone()
    sth = torch.zeros(Be, dtype=torch.int32).pin_memory()

    opt_host = _lib.BioncDynOptions()
    opt_host.fext = opt.fext
    extra_h2d = 0
    if ip_host is not None:  # the host entry point takes the per-system parameters from host memory and uploads them
        iph = torch.from_numpy(np.ascontiguousarray(ip_host[:Be])).pin_memory()
        keep.append(iph)
        opt_host.inertia = iph.data_ptr()
        extra_h2d = int(iph.numel() * 8)

    def e2e_call():
        rc = lib.bionc_cuda_rk4_host(model._h, Qh.data_ptr(), Qdh.data_ptr(), h_step, None, S, 1, C.byref(opt_host), sth.data_ptr(), Be)
        _lib.check(rc, "rk4_host")

    e2e_call()  # warm-up: allocates the library's device scratch
    barrier()
    e2e_times = []
    for _ in range(max(2, min(args.steps, 5))):
        Qh.copy_(Qh0), Qdh.copy_(Qdh0)
        t0 = time.perf_counter()
        e2e_call()
        e2e_times.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_times))
    barrier()

    # ---- N > 1: the configuration's own scaling run (1e6 systems IN TOTAL) and the final gather
    from bionc_b200.utils.sharding import gather_states, max_over_ranks

    strong_ms = gather_ms = 0.0
    gather_bytes = 0
    Bs = max(1, args.batch // world)
    if distributed:
        qs, qds = Q[:Bs], Qd[:Bs]  # views of the resident arrays: the first B / N systems of this rank
        launch(opt, qs, qds, Bs)
        strong_ms, _, _ = timed(opt, qs, qds, Bs)
        # the only collective of the workload: one all_gather_into_tensor of the final states [Q | Qdot] of the strong run
        y = torch.cat([qs, qds], dim=1).contiguous()
        full = torch.empty((world * Bs, 2 * n), dtype=torch.float64, device=dev)  # equal shards: rank r's rows are [r Bs, (r + 1) Bs)
        dist.all_gather_into_tensor(full, y)  # warm-up (NCCL channel set-up)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        dist.all_gather_into_tensor(full, y)
        g1.record(stream)
        barrier()
        gather_ms = g0.elapsed_time(g1)
        gather_bytes = int(full.numel() * 8)
        # bitwise: the few systems the unstabilised run has driven non-finite (see `flagged`) are NaN, and NaN != NaN
        bits = lambda a: a.contiguous().view(torch.int64)
        assert torch.equal(bits(full[rank * Bs : (rank + 1) * Bs]), bits(y)), "all_gather: own shard differs"
        assert torch.equal(bits(gather_states(y, world * Bs)), bits(full)), "gather_states differs from all_gather_into_tensor"
        del full, y

    t = max_over_ranks([total_ms, e2e_s, kernel_ms_mean, float(bad), stab_total_ms, float(stab_bad), strong_ms, gather_ms], device=dev)
    if distributed:
        # flagged systems are summed over the ranks, times are the slowest rank's
        cnt = torch.tensor([float(bad), float(stab_bad)], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt)
        t[3], t[5] = float(cnt[0].item()), float(cnt[1].item())
    total_ms, e2e_s, kernel_ms_mean, bad, stab_total_ms, stab_bad, strong_ms, gather_ms = t

    if rank == 0:
        peaks, peak_src = measured_peaks()
        best, sustained = C.c_double(), C.c_double()
        _lib.check(lib.bionc_cuda_measure_fp64_peak(local_rank, 10, C.byref(best), C.byref(sustained)), "measure_fp64_peak")
        value = world * B * S * args.steps / (total_ms * 1e-3)
        per_gpu_rate = B * S / (kernel_ms_mean * 1e-3)  # dominant kernel, one launch, one GPU
        fl, by = rf.flops_per_step(model.table), rf.bytes_per_step(model.table)
        lay = model.kernel_layout()
        general = fx is not None or any(int(tt) != 0 for tt in model.table.blk_type)
        kernel_name = "rk4_kernel<%s, %d threads, %d systems per CTA>" % (
            "general" if general else "lean", 32 * model.nb_segments, lay["systems_per_cta"])
        if spec_seconds is not None:
            kernel_name += ", model-specialised build (NVRTC, %.1f s once per model, outside the timed region)" % spec_seconds
        algorithmic_tf = per_gpu_rate * fl / 1e12
        ncu = NCU.get(args.workload)
        if ncu is not None and (ncu.get("build", "generic") == "model-specialised") != (spec_seconds is not None):
            ncu = None  # the committed capture is of the other build of the kernel
        same_cfg = ncu is not None and (ncu["systems"], ncu["rk4_steps_per_launch"]) == (B, S)
        executed_flop = executed_tf = frac = traffic = None
        if ncu is not None:  # per system-step figures scale to any batch; the DRAM traffic only holds at the captured configuration
            executed_flop = (2.0 * ncu["dfma"] + ncu["dmul"] + ncu["dadd"]) / (ncu["systems"] * ncu["rk4_steps_per_launch"])
            executed_tf = per_gpu_rate * executed_flop / 1e12
            frac = executed_tf / sustained.value
            traffic = ncu["dram_bytes"] if same_cfg else None
        alg_bytes = B * by  # one launch: the state is read once and written once, whatever the number of steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl["desc"], "systems_per_gpu": B, "rk4_steps_per_launch": S, "h": h_step, "normalize": True,
                "integration": "%d steps from the initial states, then reset (untimed device copy): %d launches per integration" % (
                    args.run_steps, max(1, args.run_steps // S)),
                "states": "%d pairwise distinct seeded consistent states per GPU (SURVEY 8(d) generator)" % B,
                "l2": "state per GPU = %.0f MB, larger than the 126 MB L2; every launch re-reads it from HBM" % (B * 2 * n * 8 / 1e6),
                "parallelism": f"batch sharded over {world} GPU(s), no collective on the step path",
            },
            "roofline": {
                "bound": "fp64", "achieved": executed_tf, "peak": sustained.value, "unit": "TFLOP/s", "frac": frac,
                "executed_flop_per_system_step": executed_flop,
                "fp64_pipe_active_pct_ncu": ncu.get("fp64_pipe_pct") if ncu else None,
                "counts_source": ncu.get("source") if ncu else "no committed ncu capture for this workload",
                "algorithmic_achieved": algorithmic_tf, "algorithmic_frac": algorithmic_tf / sustained.value,
                "flops_per_system_step": fl,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "traffic_ratio": (traffic / alg_bytes) if traffic else None,
                "bytes_per_system_step": by, "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peak_src,
                "dram_gbs": (traffic / (kernel_ms_mean * 1e-3) / 1e9) if traffic else None,
                "kernel": kernel_name, "kernel_ms": kernel_ms_mean,
                "peak_source": "DFMA micro-benchmark in this run (bionc_cuda_measure_fp64_peak): sustained %.2f, best launch %.2f TFLOP/s"
                               % (sustained.value, best.value),
            },
            "e2e": {
                "value": world * Be * S / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(2 * Be * n * 8) + extra_h2d,
                "d2h_bytes_per_step": int(2 * Be * n * 8 + Be * 4), "systems": Be,
                "api": "bionc_cuda_rk4_host (C ABI, pinned host buffers, H2D + kernel + D2H timed)",
            },
            "gpu_launches": gpu_launches,
            "clocks": clocks.summary(t_wall0, t_wall1),
            "flagged": {
                "systems": int(bad), "of": world * B,
                "meaning": "status != 0 after %d unstabilised steps: the drift of the rigid-body and joint constraints makes the "
                           "joint-level matrix of these free-falling chains singular or the state non-finite (the reference raises "
                           "LinAlgError or returns NaN there too); flagged systems keep computing, nothing is skipped" % args.run_steps,
            },
            "stabilized": {
                "value": world * B * S * stab_steps / (stab_total_ms * 1e-3), "unit": UNIT, "alpha": STAB["alpha"], "beta": STAB["beta"],
                "flagged_systems": int(stab_bad), "kernel": kernel_name.replace("<lean", "<lean + Baumgarte"),
            },
            "wall_s": t_wall,
        }
        if distributed:
            line["strong"] = {"value": world * Bs * S * args.steps / (strong_ms * 1e-3), "unit": UNIT, "systems_total": world * Bs,
                              "systems_per_gpu": Bs, "ms_per_step": strong_ms / args.steps}
            line["gather"] = {"ms": gather_ms, "bytes": gather_bytes, "gbs": gather_bytes / (gather_ms * 1e-3) / 1e9 if gather_ms else None,
                              "what": "all_gather_into_tensor (NCCL) of the final [Q | Qdot] of the strong run: every rank receives all systems"}
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=1_000_000, help="systems per GPU")
    ap.add_argument("--rk4-steps-per-launch", type=int, default=100)
    ap.add_argument("--run-steps", type=int, default=1000, help="RK4 steps of one integration (BASELINE.json: 1000); the states are reset after it")
    ap.add_argument("--e2e-batch", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-specialize", action="store_true", help="general-path workloads: the generic kernel instead of the model-specialised build")
    ap.add_argument("--workload", default="lower_limb", choices=sorted(WORKLOADS),
                    help="default: the configuration BASELINE.json's metric is quoted on; the others are its remaining GPU configs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
