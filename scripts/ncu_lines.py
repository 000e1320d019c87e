#!/usr/bin/env python
"""Per-CUDA-source-line stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
usage: scripts/ncu_lines.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = ''; out = []; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8 or r[0] == '': continue
    try:
        out.append((int(r[4]), int(r[7]), fname, r[0], r[1].strip()[:120]))
    except ValueError:
        pass
tot = sum(o[0] for o in out)
print('total samples', tot)
for w, ins, f, ln, src in sorted(out, key=lambda x: -x[0])[:top]:
    print('%6d %5.1f%% inst=%9d %s:%s | %s' % (w, 100.0 * w / tot, ins, f, ln, src))
