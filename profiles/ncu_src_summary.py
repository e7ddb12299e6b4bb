"""Summarise an `ncu --page source --csv` dump: executed instructions and stall samples per opcode,
and the hottest SASS lines.  Usage: python profiles/ncu_src_summary.py prof_src.csv [top]"""
import csv
import collections
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0, 0.0])
lines = []
tot_inst = tot_samp = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    try:
        inst = int(float(r[col["Instructions Executed"]] or 0))
        samp = int(float(r[col["# Samples"]] or 0))
        thr = float(r[col["Avg. Threads Executed"]] or 0)
    except ValueError:
        continue
    toks = src.split()
    op = toks[0]
    if op.startswith("@"):
        op = toks[1] if len(toks) > 1 else op
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST")) and "." in op else "")
    ops[op][0] += inst
    ops[op][1] += samp
    ops[op][2] += inst * thr
    tot_inst += inst
    tot_samp += samp
    lines.append((samp, inst, src[:90]))
print(f"total instructions {tot_inst:,}  samples {tot_samp:,}")
print(f"{'opcode':14s} {'inst %':>7s} {'samples %':>9s} {'avg thr':>7s}")
for op, (i, s, t) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{op:14s} {100*i/tot_inst:7.2f} {100*s/max(1,tot_samp):9.2f} {t/max(1,i):7.1f}")
print("\nhottest lines by stall samples")
for s, i, src in sorted(lines, reverse=True)[:top]:
    print(f"{100*s/max(1,tot_samp):6.2f}%  {i:>10,}  {src}")
