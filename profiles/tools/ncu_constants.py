#!/usr/bin/env python
"""From the `ncu --set full` captures of the fused integrator (one per workload, gpurun_out/r02_ncu_<workload>.ncu-rep) to
  * profiles/ncu_constants.json   what bench.py's roofline needs: executed FP64 thread instructions, all warp instructions,
                                  FP64 pipe utilisation, DRAM bytes, per launch at the captured configuration
  * profiles/r02_ncu_<workload>.txt   the key-metric summary (profiles/ncu_raw_summary.py) kept for the record
Run here (no GPU needed):  python profiles/tools/ncu_constants.py [version tag]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = {}
for w in ("lower_limb", "pendulum_force", "knee", "full_body", "n_link_4"):
    rawcsv = os.path.join(REPO, "gpurun_out", f"r02_ncu_{w}_raw.csv")  # written on the GPU box by profiles/tools/ncu_all.sh
    if not os.path.exists(rawcsv) or os.path.getsize(rawcsv) < 1000:
        continue
    raw = open(rawcsv).read()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: v for h, v in zip(hdr, vals)}
    u = {h: x for h, x in zip(hdr, units)}

    def val(key):
        x = float(d[key])
        unit = u[key]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0}.get(unit, 1.0)
        return x * scale

    def fp64(op):  # the raw page lists the per-cycle rate of the sum over the SM sub-partitions
        return float(d[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"]) * float(d["smsp__cycles_elapsed.avg"])

    csvp = os.path.join(REPO, "scratch", f"{TAG}_{w}_raw.csv")
    os.makedirs(os.path.dirname(csvp), exist_ok=True)
    open(csvp, "w").write(raw)
    summary = subprocess.run([sys.executable, os.path.join(REPO, "profiles", "ncu_raw_summary.py"), csvp], capture_output=True, text=True).stdout
    grid, spc_line = int(float(d["launch__grid_size"])), d["Kernel Name"]
    # systems of the captured launch: CTAs x systems per CTA (4th template argument); the launch log says which build ran
    targs = [int(x) for x in re.findall(r"\d+", spc_line[spc_line.index("<"):spc_line.index(">")])]
    systems = min(1_000_000, grid * targs[3])
    logp = os.path.join(REPO, "gpurun_out", f"r02_launch_{w}.log")
    specialised = os.path.exists(logp) and "specialised in" in open(logp).read()
    txt = os.path.join(REPO, "profiles", f"{TAG}_ncu_{w}.txt")
    with open(txt, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch of the fused integrator, bench.py --workload {w} ({systems} systems x 100 steps)"
                f"{', model-specialised build of the kernel (bionc_cuda_model_specialize)' if specialised else ''}\n")
        f.write(f"# kernel: {spc_line}\n")
        f.write(summary)
        for op in ("dfma", "dmul", "dadd"):
            f.write(f"{'smsp__sass_thread_inst_executed_op_' + op + '_pred_on.sum  (= .sum.per_cycle_elapsed x smsp__cycles_elapsed.avg)':110s} {fp64(op):.6e} inst\n")
    out[w] = {
        "systems": systems, "rk4_steps_per_launch": 100, "build": "model-specialised" if specialised else "generic",
        "dfma": fp64("dfma"), "dmul": fp64("dmul"), "dadd": fp64("dadd"),
        "warp_instructions": val("smsp__inst_executed.sum"),
        "fp64_pipe_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
        "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "kernel_ms_under_ncu": val("gpu__time_duration.sum") * 1e3,
        "registers": int(float(d["launch__registers_per_thread"])),
        "kernel": d["Kernel Name"], "source": f"profiles/{TAG}_ncu_{w}.txt",
    }
    flop = (2 * out[w]["dfma"] + out[w]["dmul"] + out[w]["dadd"]) / (systems * 100.0)
    print(f"{w:16s} {out[w]['kernel_ms_under_ncu']:9.2f} ms  fp64 pipe {out[w]['fp64_pipe_pct']:5.1f} %  issue {out[w]['issue_active_pct']:5.1f} %  "
          f"executed flop/system-step {flop:9.0f}  dram {out[w]['dram_bytes'] / 1e9:7.2f} GB  regs {out[w]['registers']}")
with open(os.path.join(REPO, "profiles", "ncu_constants.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
