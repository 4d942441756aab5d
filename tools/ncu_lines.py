#!/usr/bin/env python3
"""Top source lines by warp-stall samples from an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
sec = rows[hdrs[0] + 1: hdrs[1] if len(hdrs) > 1 else len(rows)]
def _i(x):
    try:
        return int(x)
    except ValueError:
        return 0
src = [(int(r[0]), r[1], _i(r[6]), _i(r[7])) for r in sec if r[0].isdigit()]
tot = sum(s[2] for s in src) or 1
print('total samples', tot)
for ln, code, smp, ins in sorted(src, key=lambda x: -x[2])[:top]:
    print(f"{ln:5d} {smp:6d} ({100*smp/tot:4.1f}%) inst={ins:9d}  {code.strip()[:110]}")
