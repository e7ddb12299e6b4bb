"""Print the JSON lines of config_sweep.py / post_sweep.py / bench.py compactly.  python profiles/tools/show_sweep.py file ..."""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        l = l.strip()
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        if "evaluator" in d:
            print(f"{d['evaluator']:55s} B={d['systems']:8d} {d['ms']:8.3f} ms {d['GB_per_s']:8.0f} GB/s {d['frac_of_hbm_peak']:.2f}")
        elif "system_steps_per_s" in d:
            print(f"{d['model']:22s} N={d['segments']:2d} {d['system_steps_per_s']:.3e} steps/s {d['kernel_ms']:8.2f} ms flagged {d['flagged']} "
                  f"spc {d['layout']['systems_per_cta']} smem {d['layout']['smem_bytes_per_cta']} schedule {d['layout']['serial_joint_level_schedule']}")
        elif "metric" in d:
            r = d.get("roofline", {})
            print(f"{d['config']['workload'][:40]:40s} n_gpus {d['n_gpus']} value {d['value']:.3e} e2e {d['e2e']['value']:.3e} kernel_ms {r.get('kernel_ms')} "
                  f"frac {r.get('frac')} alg_frac {r.get('algorithmic_frac')} stab {d.get('stabilized', {}).get('value')} flagged {d.get('flagged', {}).get('systems')} "
                  f"cpu {d.get('cpu_baseline', {}).get('value')} {d.get('cpu_baseline', {}).get('kind')} strong {d.get('strong', {}).get('value')} gather_ms {d.get('gather', {}).get('ms')}")
