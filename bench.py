#!/usr/bin/env python
"""bench.py -- RK4 forward-dynamics throughput of the bionc_cuda backend (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config 3 of BASELINE.json): the 4-segment lower-limb model (PELVIS free, hip / knee /
ankle spherical; geometry of the reference's examples/models/lower_limb.nc, inertia of SURVEY.md
8(d)), B = 1e6 independent systems PER GPU with seeded, constraint-consistent random initial states,
h = 0.01 s, post-step renormalisation on.  One bench "step" = one launch of the fused integrator
advancing every system by `rk4_steps_per_launch` (default 100) RK4 steps, 4 forward-dynamics evaluations
each -- BASELINE.json's configurations integrate 1000 steps, i.e. ten such launches.
Metric: RK4 system-steps/s, whole job (all GPUs).

  value        states resident in HBM, CUDA events on the launching stream, max over ranks
  e2e          the same through the C ABI's host-buffer entry point (bionc_cuda_rk4_host): pinned
               host arrays in, H2D + kernel + D2H inside the timed region
  roofline     algorithmic FLOPs (SURVEY.md 8(d) count) / measured kernel time vs the FP64 DFMA peak
               measured in this run, plus the HBM fraction against MEASURED_PEAKS.json
  cpu_baseline the numpy oracle port of the reference algorithm on this box's host cores (N=1, rank 0)

`--impl reference` times the reference's CPU algorithm (oracle port; the reference itself is pure
Python + numpy and does not travel to the GPU box) on all host cores.
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

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

MODEL = os.path.join(REPO, "tests", "golden", "models", "lower_limb.npz")
WORKLOAD = "lower_limb_4seg (BASELINE.json configs[2]; PELVIS free + 3 spherical joints, n=48, m=33)"
H = 0.01
# The default workload is the one BASELINE.json's metric is quoted on (the 4-segment lower limb).  `--workload`
# runs the other BASELINE configurations through the same measurement (SURVEY 8(d) definitions); they are not
# bench lines of the driver, which calls bench.py without that flag.
WORKLOADS = {
    "lower_limb": dict(model="lower_limb", states="tree", h=0.01, fext=False, inertia=False, desc=WORKLOAD),
    "pendulum_force": dict(model="pendulum_force_batch", states="hinge", h=1.0 / 60.0, fext=True, inertia=False,
                           desc="pendulum_with_force.nmod + no_equilibrium force (BASELINE.json configs[1]; ground hinge, n=12, m=11)"),
    "knee": dict(model="knee_feikes", states="knee", h=5e-5, fext=False, inertia=False,
                 desc="knee_feikes parallel mechanism (BASELINE.json configs[3]; closed loops, n=24, m=20)"),
    "full_body": dict(model="full_body", states="tree", h=0.001, fext=False, inertia=True,
                      desc="synthetic full body, per-system inertial parameters (BASELINE.json configs[4]; 10 segments, n=120, m=87)"),
}


def workload_paths(name):
    w = WORKLOADS[name]
    return (os.path.join(REPO, "tests", "golden", "models", w["model"] + ".npz"),
            os.path.join(REPO, "tests", "golden", "cases", w["model"] + ".npz"))


def workload_states(name, table, count, seed):
    """Seeded consistent initial states of a workload (numpy, SURVEY 8(d) generators)."""
    from bionc_b200.utils import states as st

    kind = WORKLOADS[name]["states"]
    if kind == "tree":
        return st.tree_states(table, count, seed=seed)
    if kind == "hinge":
        return st.hinge_pendulum_states(count, seed=seed)
    with np.load(workload_paths(name)[1]) as z:  # knee: rigid motions of the reference's own initial state
        q0 = np.array(z["Q"][0])
    return st.rigidly_moved_states(q0, count, seed=seed, angle_max=0.3, omega_sigma=0.0)


def workload_fext(name):
    """(segment, kind, param) arrays of the workload's external forces (from the golden case), or None."""
    if not WORKLOADS[name]["fext"]:
        return None
    with np.load(workload_paths(name)[1]) as z:
        return np.array(z["fext_seg"]), np.array(z["fext_kind"]), np.array(z["fext_param"])
# DRAM traffic of ONE launch of the dominant kernel at the default configuration, from the committed
# `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum): 0.79 GB read (the 768 MB of
# state, once) + 8.9 GB written (768 MB of state + register spills evicted from L2: the kernel runs at 128
# registers/thread to keep 4 CTAs per SM resident; before y0 moved to tensor memory the same launch wrote
# 66 GB).  The algorithmic bytes of the same launch are B * S * 1536 B = 153.6 GB: the state stays on chip
# for the S steps of a launch; DRAM runs below 1 % of peak.
NCU_TRAFFIC = {"systems_per_gpu": 1_000_000, "rk4_steps_per_launch": 100, "bytes": 792.342528e6 + 8.927182e9,
               "source": "profiles/r01_ncu_rk4_kernel_v12_benchconfig.txt"}
METRIC = "rk4_forward_dynamics_system_steps_per_s"
UNIT = "system-steps/s"


# ----------------------------------------------------------------------------- CPU side (oracle port)
def _cpu_worker(args):
    """One host core: integrate `nsys` systems for `nsteps` RK4 steps with the oracle port."""
    seed, nsys, nsteps, workload = args
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from bionc_oracle import OracleModel  # noqa: E402  (the CPU baseline leg is allowed to run the oracle)

    from bionc_b200.bionc_cuda.model_table import ModelTable
    from bionc_b200.utils import states as st

    model_file = workload_paths(workload)[0]
    table = ModelTable.load(model_file)
    om = OracleModel(model_file)
    Q, Qd = workload_states(workload, table, nsys, seed)
    fx = workload_fext(workload)
    fext = None if fx is None else [(int(a), int(b), c) for a, b, c in zip(*fx)]
    per_system = None
    if WORKLOADS[workload]["inertia"]:
        per_system = st.perturbed_inertial_parameters(table, nsys, seed=seed)
    t = np.arange(nsteps + 1) * WORKLOADS[workload]["h"]
    t0 = time.perf_counter()
    for s in range(nsys):
        if per_system is not None:  # like the reference, one model object per parameter set (a few ms, inside the timing)
            om = OracleModel(model_file, seg_mass=per_system[0][s], seg_com=per_system[1][s], seg_J=per_system[2][s])
        om.rk4(Q[s], Qd[s], t, normalize=True, fext=fext)
    return time.perf_counter() - t0


def cpu_rate(nsys_per_core: int, nsteps: int, cores: int, seed0: int = 1000, workload: str = "lower_limb"):
    """Aggregate system-steps/s of `cores` worker processes, wall clock of the slowest worker."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, [(seed0 + c, nsys_per_core, nsteps, workload) for c in range(cores)])
    wall = max(times)  # integration time of the slowest worker (process start-up and imports excluded)
    return cores * nsys_per_core * nsteps / wall, wall


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm on all host cores.  Each bench step is a bounded
    sample: every core integrates `nsys` lower-limb systems for `nsteps` RK4 steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    nsys, nsteps = 4, 100  # ~1 s per core per bench step at a few hundred system-steps/s/core
    for _ in range(args.warmup):
        cpu_rate(nsys, nsteps, cores, workload=args.workload)
    wall = 0.0
    for k in range(args.steps):
        wall += cpu_rate(nsys, nsteps, cores, seed0=2000 + 100 * k, workload=args.workload)[1]
    value = args.steps * cores * nsys * nsteps / wall
    sample = f"{cores} cores x {nsys} systems x {nsteps} RK4 steps per bench step (oracle port, numpy dense KKT LU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["desc"], "h": WORKLOADS[args.workload]["h"], "normalize": True},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
        nsys, nsteps = 8, 1000  # ~10-25 s of CPU work per core at a few hundred system-steps/s/core
        if WORKLOADS[args.workload]['model'] == 'full_body':
            nsys, nsteps = 4, 250  # 2.4 ms per evaluation
        rate, wall = cpu_rate(nsys, nsteps, cores, workload=args.workload)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{cores} cores x {nsys} systems x {nsteps} RK4 steps, same model/h/normalize ({wall:.1f} s wall)"}

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
    # seeded synthetic initial states: a pool of 2^16 consistent states (numpy), tiled to B on the device
    pool = min(B, 1 << 16)
    Qp, Qdp = workload_states(args.workload, model.table, pool, seed=1 + rank)
    reps = (B + pool - 1) // pool
    Q0 = torch.from_numpy(Qp).to(dev).repeat(reps, 1)[:B].contiguous()
    Qd0 = torch.from_numpy(Qdp).to(dev).repeat(reps, 1)[:B].contiguous()
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
    ip_pool = None
    if wl["inertia"]:  # config 5: per-system natural inertial parameters (B, N, 10), resident on the device
        mass, com, J = st.perturbed_inertial_parameters(model.table, pool, seed=100 + rank)
        ip_pool = np.ascontiguousarray(np.concatenate(
            [mass[..., None], com, np.stack([J[..., 0, 0], J[..., 0, 1], J[..., 0, 2], J[..., 1, 1], J[..., 1, 2], J[..., 2, 2]], axis=-1)], axis=-1))
        ip_dev = torch.from_numpy(ip_pool).to(dev).repeat(reps, 1, 1)[:B].contiguous()
        keep.append(ip_dev)
        opt.inertia = ip_dev.data_ptr()
    stream = torch.cuda.current_stream(dev)

    def launch():
        rc = lib.bionc_cuda_rk4(model._h, Q.data_ptr(), Qd.data_ptr(), h_step, None, S, 1, C.byref(opt), None, 1, None,
                                status.data_ptr(), B, C.c_void_p(stream.cuda_stream))
        _lib.check(rc, "rk4")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(args.warmup):
        launch()
    barrier()
    # every timed launch starts from the same consistent initial states: a free-falling chain integrated
    # for thousands of steps drifts off to large coordinates; resetting keeps the timed work representative
    Q.copy_(Q0), Qd.copy_(Qd0)
    launches0 = lib.bionc_cuda_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        e0.record(stream)
        launch()
        e1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0
    time.sleep(0.1)  # let the last samples of the window arrive
    clocks.__exit__()
    gpu_launches = int(lib.bionc_cuda_launch_count() - launches0)
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = evs[0][0].elapsed_time(evs[-1][1])  # first launch start -> last launch end, on the device
    bad = int((status != 0).sum().item())

    # ---- end-to-end through the C ABI with host buffers (pinned), copies inside the timed region
    Be = min(B, args.e2e_batch)
    Qh = torch.from_numpy(np.ascontiguousarray(Qp[:Be] if Be <= pool else np.tile(Qp, (reps, 1))[:Be])).pin_memory()
    Qdh = torch.from_numpy(np.ascontiguousarray(Qdp[:Be] if Be <= pool else np.tile(Qdp, (reps, 1))[:Be])).pin_memory()
    Qh0, Qdh0 = Qh.clone(), Qdh.clone()
    sth = torch.zeros(Be, dtype=torch.int32).pin_memory()

    opt_host = _lib.BioncDynOptions()
    opt_host.fext = opt.fext
    extra_h2d = 0
    if ip_pool is not None:  # the host entry point takes the per-system parameters from host memory and uploads them
        iph = torch.from_numpy(np.ascontiguousarray(ip_pool[:Be] if Be <= pool else np.tile(ip_pool, (reps, 1, 1))[:Be])).pin_memory()
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

    # ---- reduce over ranks: slowest rank defines the job time
    from bionc_b200.utils.sharding import gather_states, max_over_ranks

    t = max_over_ranks([total_ms, e2e_s, float(np.mean(kernel_ms)), float(bad)], device=dev)
    if distributed:
        # the only collective of the workload: a final gather (of per-rank checksums by default;
        # --gather-states gathers the full states, 768 MB per rank at the default batch)
        ok = (status == 0).to(torch.float64)[:, None]  # a handful of systems diverge in 800+ unstabilised steps, as the reference does
        chk = torch.stack([torch.nan_to_num(Q * ok).sum(), torch.nan_to_num(Qd * ok).sum()])[None]
        allchk = gather_states(chk, world)
        assert allchk.shape == (world, 2) and bool(torch.isfinite(allchk).all())
        if args.gather_states:
            full = gather_states(Q, world * B)
            assert full.shape == (world * B, n)
    total_ms, e2e_s, kernel_ms_mean, bad = t

    if rank == 0:
        peaks, peak_src = measured_peaks()
        best, sustained = C.c_double(), C.c_double()
        _lib.check(lib.bionc_cuda_measure_fp64_peak(local_rank, 10, C.byref(best), C.byref(sustained)), "measure_fp64_peak")
        value = world * B * S * args.steps / (total_ms * 1e-3)
        per_gpu_rate = B * S / (kernel_ms_mean * 1e-3)  # dominant kernel, one launch, one GPU
        fl, by = rf.flops_per_step(model.table), rf.bytes_per_step(model.table)
        lay = model.kernel_layout()
        general = fx is not None or any(int(t) != 0 for t in model.table.blk_type)
        kernel_name = "rk4_kernel<%s, %d threads, %d systems per CTA%s>" % (
            "general" if general else "lean", 32 * model.nb_segments, lay["systems_per_cta"],
            ", y0 in tensor memory" if model.nb_segments >= 3 else "")
        achieved_tf = per_gpu_rate * fl / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl["desc"], "systems_per_gpu": B, "rk4_steps_per_launch": S, "h": h_step, "normalize": True,
                "l2": "state per GPU = %.0f MB, larger than the 126 MB L2; every launch re-reads it from HBM" % (B * 2 * n * 8 / 1e6),
                "parallelism": f"batch sharded over {world} GPU(s), no collective on the step path",
            },
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": sustained.value, "unit": "TFLOP/s",
                "frac": achieved_tf / sustained.value,
                "traffic": NCU_TRAFFIC["bytes"] if (args.workload, B, S) == ("lower_limb", NCU_TRAFFIC["systems_per_gpu"], NCU_TRAFFIC["rk4_steps_per_launch"]) else None,
                "traffic_source": NCU_TRAFFIC["source"], "algorithmic_bytes_per_launch": B * S * by,
                "kernel": kernel_name, "kernel_ms": kernel_ms_mean,
                "flops_per_system_step": fl, "bytes_per_system_step": by,
                "peak_source": "DFMA micro-benchmark in this run (bionc_cuda_measure_fp64_peak): sustained %.2f, best launch %.2f TFLOP/s"
                               % (sustained.value, best.value),
                "hbm": {"achieved": per_gpu_rate * by / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": per_gpu_rate * by / 1e9 / peaks.get("hbm_gbs", 6650.0), "peak_source": peak_src},
            },
            "e2e": {
                "value": world * Be * S / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(2 * Be * n * 8) + extra_h2d,
                "d2h_bytes_per_step": int(2 * Be * n * 8 + Be * 4), "systems": Be,
                "api": "bionc_cuda_rk4_host (C ABI, pinned host buffers, H2D + kernel + D2H timed)",
            },
            "gpu_launches": gpu_launches,
            "clocks": clocks.summary(t_wall0, t_wall1),
            "status_flagged_systems": int(bad),
            "wall_s": t_wall,
        }
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
    ap.add_argument("--e2e-batch", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-states", action="store_true")
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
