"""Refresh the measured tables of DESIGN.md (between <!-- NAME --> / <!-- /NAME --> markers) from profiles/:
python profiles/tools/fill_design.py"""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = subprocess.run([sys.executable, os.path.join(REPO, "profiles", "tools", "make_tables.py")], capture_output=True, text=True, check=True).stdout
roof, evalt = out.strip().split("\n\n")
path = os.path.join(REPO, "DESIGN.md")
s = open(path).read()
for name, body in (("ROOFLINE_TABLE", roof), ("EVALUATOR_TABLE", evalt)):
    block = f"<!-- {name} -->\n{body}\n<!-- /{name} -->"
    if f"<!-- {name} -->" in s:
        s = re.sub(rf"<!-- {name} -->.*?<!-- /{name} -->", lambda m: block, s, flags=re.S)
    else:
        s = s.replace(name, block, 1)
open(path, "w").write(s)
