#!/usr/bin/env python
"""Per-CUDA-source-line sample totals of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out = []
fname = ""
h = None
for x in rows:
    if x and x[0] == "File Path": fname = x[1].split("/")[-1]
    elif x and x[0] == "Line No": h = x
    elif h and len(x) == len(h) and x[0].isdigit():
        isamp = h.index('# Samples'); iex = h.index('Instructions Executed')
        st = {h[i]: int(x[i] or 0) for i in range(len(h)) if h[i].startswith('stall_') and x[i] not in ('', '-') and int(x[i] or 0) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        out.append((int(x[isamp] or 0), int(x[iex] or 0), fname, x[0], x[1].strip()[:90], st))
tot = sum(o[0] for o in out)
print("total samples", tot)
for o in sorted(out, key=lambda o: -o[0])[:N]:
    print(f"{o[0]:8d} {100*o[0]/tot:5.1f}% {o[1]:12d} {o[2]}:{o[3]:>4s} {o[4]:90s} {o[5]}")
