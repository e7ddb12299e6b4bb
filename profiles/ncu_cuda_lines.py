"""Hot CUDA source lines of an ncu report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv;
python profiles/ncu_cuda_lines.py f.csv [top]`.  Rows with an Address of "-" are the per-source-line aggregates."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and cur and len(r) == len(hdr) and r[2] == "-":
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try:
            out.append((float(r[i_s] or 0), float(r[i_i] or 0), cur, r[0], r[1].strip()[:120]))
        except ValueError:
            pass
tot, toti = sum(o[0] for o in out) or 1.0, sum(o[1] for o in out) or 1.0
print("samples %.0f  warp instructions %.0f" % (tot, toti))
print("samples%  inst%   file:line  source")
for o in sorted(out, reverse=True)[:top]:
    print("%6.2f %6.2f  %s:%s  %s" % (100 * o[0] / tot, 100 * o[1] / toti, o[2], o[3], o[4]))
