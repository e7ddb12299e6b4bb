"""Markdown tables of DESIGN.md from the committed measurement files:  python profiles/tools/make_tables.py"""
import json
import os

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = os.path.join(REPO, "profiles")
ncu = json.load(open(os.path.join(P, "ncu_constants.json")))
R1 = {"lower_limb": 4.44e8, "pendulum_force": 7.4e8, "knee": 1.89e8, "full_body": 7.3e7, "n_link_4": 4.2e7}
print("| configuration (1e6 systems, 100 steps per launch, 1 × B200) | system-steps/s (round 1) | end to end | executed FP64 flop / system-step | **executed frac** of the DFMA peak | ncu FP64 pipe / issue | algorithmic frac | DRAM traffic ÷ algorithmic bytes | stabilised | CPU reference (cores) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for w in ("lower_limb", "pendulum_force", "knee", "full_body", "n_link_4"):
    f = os.path.join(P, f"r02_bench_{w}.json")
    if not os.path.exists(f):
        continue
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    r, c = d["roofline"], d.get("cpu_baseline", {})
    n = ncu.get(w, {})
    print(f"| {d['config']['workload'].split(' (')[0]} | **{d['value']:.2e}** ({R1[w]:.2e}) | {d['e2e']['value']:.2e} | {r['executed_flop_per_system_step']:.0f} | "
          f"**{r['frac']:.2f}** | {n.get('fp64_pipe_pct', 0):.0f} % / {n.get('issue_active_pct', 0):.0f} % | {r['algorithmic_frac']:.2f} | "
          f"{(r['traffic_ratio'] or 0):.2f} | {d['stabilized']['value']:.2e} | {c.get('value', 0):.0f} ({c.get('cores')}, {c.get('kind')}) |")
print()
f = os.path.join(P, "r02_post_sweep.json")
if os.path.exists(f):
    print("| evaluator | systems | ms | GB/s (algorithmic bytes) | of the measured HBM peak |")
    print("|---|---|---|---|---|")
    for l in open(f):
        d = json.loads(l)
        print(f"| `{d['evaluator']}` | {d['systems']:.0e} | {d['ms']:.3f} | {d['GB_per_s']:.0f} | {d['frac_of_hbm_peak']:.2f} |")
