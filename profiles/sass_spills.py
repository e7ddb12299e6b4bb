#!/usr/bin/env python
"""Where are the local-memory (spill) and FP64 instructions of a kernel?  Static SASS statistics per source line.

    nvdisasm --print-line-info-inline file.cubin > all.sass
    python profiles/sass_spills.py all.sass <substring of the mangled kernel name>

Prints the instruction count, the opcode mix (DFMA / DMUL / DADD / LDS / STS / LDL / STL / BAR ...) and the source
lines holding the local loads / stores, so that register-pressure work can be aimed before GPU time is spent."""
import collections
import re
import sys


def main():
    path, key = sys.argv[1], sys.argv[2]
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and key in l)
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//---------------------")), len(lines))
    cur = None
    chain = []
    ops = collections.Counter()
    where = collections.Counter()
    prev_was_marker = False
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if not prev_was_marker:
                chain = []
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
            prev_was_marker = True
            # attribute to the outermost frame inside dynamics.cuh (the phase of the evaluation), else the outermost frame
            dyn = [c for c in chain if c[0] == "dynamics.cuh"]
            cur = dyn[-1] if dyn else chain[-1]
            continue
        prev_was_marker = False
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        op = m.group(1).split(".")[0]
        ops[op] += 1
        if op in ("LDL", "STL"):
            where[(cur, op)] += 1
    total = sum(ops.values())
    print(f"{total} instructions")
    print(", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
    fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    print(f"FP64 {fp64} ({100.0 * fp64 / total:.1f} %): DFMA {ops['DFMA']}, DMUL {ops['DMUL']}, DADD {ops['DADD']}")
    for (loc, op), n in sorted(where.items(), key=lambda kv: -kv[1])[:30]:
        print(f"  {op} x{n:3d}  {loc}")


if __name__ == "__main__":
    main()
