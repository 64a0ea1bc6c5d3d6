#!/usr/bin/env python
"""Hot SASS instructions of an ncu report in address order, with the CUDA source line of each (nvdisasm -g of lib.so).
usage: ncu_hot.py report.ncu-rep lib.so kernel_substring [min_pct]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
i0 = starts[0]; i1 = min([j for j in starts if j > i0] + [len(rows)])
hdr = rows[i0 + 1]; data = [r for r in rows[i0 + 2:i1] if r]
ia, isamp, iaddr, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address'), hdr.index('Source')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and '(' not in h]
base = int(data[0][iaddr], 16)
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = {}
for f in os.listdir(tmp):
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infun = False; line = None
    for ln in out.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+),', ln)
        if m: infun = kname in m.group(1); continue
        if not infun: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            line = f"{os.path.basename(m.group(1))}:{m.group(2)}" + (f"<{m.group(5)}" if m.group(3) else ""); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m and line: lines[int(m.group(1), 16)] = line
    if lines: break
tots = sum(float(r[isamp] or 0) for r in data)
print(f"total samples {tots:.0f}; columns: addr samples% cum% executed source-line | sass | top stall")
cum = 0.0
for r in data:
    s = float(r[isamp] or 0); cum += s
    if s / tots * 100 >= minp:
        off = int(r[iaddr], 16) - base
        st = sorted(((float(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        print(f"{off:6x} {s/tots*100:5.2f} {cum/tots*100:5.1f} {float(r[ia] or 0):10.0f} {lines.get(off, '?'):28s} | {r[isrc][:60]:60s} | " + ' '.join(f"{h[6:]}={v:.0f}" for v, h in st))
