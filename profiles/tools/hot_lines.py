"""Hot source lines of one `ncu --page source --print-source cuda,sass --csv` export:
python profiles/tools/hot_lines.py <source.csv> [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, out, total, tinst = "", [], 0, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0].isdigit() and len(r) >= 12:
        # the source text may itself contain commas: the counters start after the first "-", "-" pair (address, SASS)
        k = next((i for i in range(2, len(r) - 5) if r[i] == "-" and r[i + 1] == "-" and r[i + 2].isdigit()), -1)
        if k < 0:
            continue
        samples, inst = int(r[k + 2]), int(r[k + 5])
        out.append((samples, inst, cur, int(r[0]), ",".join(r[1:k])[:110]))
        total += samples
        tinst += inst
print(f"total samples {total}, warp instructions {tinst}")
for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100.0 * s / max(total, 1):5.2f}% smp {100.0 * i / max(tinst, 1):5.2f}% inst  {f}:{ln}  {src.strip()}")
