"""Four launches of the fused integrator at the bench configuration of a workload (1e6 systems x 100 steps), nothing else:
the cheap target for `ncu -k regex:rk4_kernel -s 3 -c 1`.  States are a pool of 2^15 tiled to the batch (the instruction
counts and DRAM traffic of a launch do not depend on the states being distinct).
    python profiles/tools/rk4_launch.py <workload> [--stab] [--spec] [--check] [--batch=N]
--spec: the model-specialised build of the kernel (bionc_cuda_model_specialize); --check: the same launch by the generic kernel,
largest difference."""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from bionc_b200.bionc_cuda import BiomechanicalModel, _lib  # noqa: E402
from bionc_b200.utils import states as st  # noqa: E402

name = sys.argv[1]
stab = "--stab" in sys.argv
B, S = 1_000_000, 100
for a in sys.argv:
    if a.startswith("--batch="):
        B = int(a[8:])
wl = bench.WORKLOADS[name]
model = BiomechanicalModel.load(bench.workload_paths(name)[0])
pool = 1 << 15
Qp, Qdp = bench.workload_states(name, model.table, pool, seed=1)
reps = (B + pool - 1) // pool
Q = torch.from_numpy(Qp).cuda().repeat(reps, 1)[:B].contiguous()
Qd = torch.from_numpy(Qdp).cuda().repeat(reps, 1)[:B].contiguous()
lib = _lib.load()
opt = _lib.BioncDynOptions()
keep = []
fx = bench.workload_fext(name)
if fx is not None:
    seg, kind, par = (np.ascontiguousarray(fx[0], dtype=np.int32), np.ascontiguousarray(fx[1], dtype=np.int32), np.ascontiguousarray(fx[2], dtype=np.float64))
    fdesc = _lib.BioncFextDesc(int(seg.size), seg.ctypes.data_as(_lib.c_int32_p), kind.ctypes.data_as(_lib.c_int32_p), par.ctypes.data_as(_lib.c_double_p))
    keep += [seg, kind, par, fdesc]
    opt.fext = C.pointer(fdesc)
if wl["inertia"]:
    mass, com, J = st.perturbed_inertial_parameters(model.table, pool, seed=100)
    ip = np.concatenate([mass[..., None], com, np.stack([J[..., 0, 0], J[..., 0, 1], J[..., 0, 2], J[..., 1, 1], J[..., 1, 2], J[..., 2, 2]], axis=-1)], axis=-1)
    ipd = torch.from_numpy(np.ascontiguousarray(ip)).cuda().repeat(reps, 1, 1)[:B].contiguous()
    keep.append(ipd)
    opt.inertia = ipd.data_ptr()
if stab:
    opt.use_stabilization, opt.alpha, opt.beta = 1, bench.STAB["alpha"], bench.STAB["beta"]
if "--spec" in sys.argv:  # this model's own build of the kernel (bionc_cuda_model_specialize)
    import time
    t0 = time.perf_counter()
    _lib.check(lib.bionc_cuda_model_specialize(model._h, C.byref(opt), 0, 1), "specialize")
    print(f"specialised in {time.perf_counter() - t0:.1f} s, kernels loaded: {lib.bionc_cuda_model_specialized(model._h, None)}")
status = torch.zeros(B, dtype=torch.int32, device="cuda")
Q0, Qd0 = Q.clone(), Qd.clone()
ev = []
for k in range(4):
    Q.copy_(Q0), Qd.copy_(Qd0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.bionc_cuda_rk4(model._h, Q.data_ptr(), Qd.data_ptr(), wl["h"], None, S, 1, C.byref(opt), None, 1, None, status.data_ptr(), B,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "rk4")
    e1.record()
    ev.append((e0, e1))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
if "--check" in sys.argv:  # the same launch by the generic kernel (a second handle of the model): largest difference
    ref = BiomechanicalModel.load(bench.workload_paths(name)[0])
    Qr, Qdr = Q0.clone(), Qd0.clone()
    _lib.check(lib.bionc_cuda_rk4(ref._h, Qr.data_ptr(), Qdr.data_ptr(), wl["h"], None, S, 1, C.byref(opt), None, 1, None, status.data_ptr(), B,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "rk4")
    torch.cuda.synchronize()
    ok = torch.isfinite(Qr).all(dim=1) & torch.isfinite(Q).all(dim=1)
    print(f"vs generic kernel after {S} steps: max |dQ| {float((Q - Qr)[ok].abs().max()):.3e}, max |dQdot| {float((Qd - Qdr)[ok].abs().max()):.3e}, "
          f"finite in both {int(ok.sum())} / {B}, bitwise equal {bool(torch.equal(Q[ok], Qr[ok]))}")
print(f"{name}{' stabilised' if stab else ''}: {ms[-1]:.2f} ms per launch, {B * S / (ms[-1] * 1e-3):.3e} system-steps/s, flagged {int((status != 0).sum())}, layout {model.kernel_layout()}")
