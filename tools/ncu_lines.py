#!/usr/bin/env python
"""Per-source-line roll-up of an ncu report: instructions executed / stall samples per CUDA source line.
usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [n_samples] [top]
Maps ncu's SASS page (per-instruction counters) onto `nvdisasm -g` line info of the same kernel in lib.so."""
import csv, io, os, re, subprocess, sys, tempfile
rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
nsamp = float(sys.argv[4]) if len(sys.argv) > 4 else None
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
# several launches may be in the report: take the LAST one whose (demangled) name contains argv[6] (default: first launch)
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
pick = sys.argv[6] if len(sys.argv) > 6 else None
cand = [i for i in starts if pick is None or pick in rows[i][1]]
i0 = cand[-1] if pick else cand[0]
i1 = min([j for j in starts if j > i0] + [len(rows)])
mangled_hint = rows[i0][1]
hdr = rows[i0 + 1]; data = [r for r in rows[i0 + 2:i1] if r]
ia, isamp, iaddr, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address'), hdr.index('Source')
base = int(data[0][iaddr], 16)
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = {}
for f in os.listdir(tmp):
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur = None; infun = False; line = None
    for ln in out.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+),', ln)
        if m:
            infun = kname in m.group(1); continue
        if not infun: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m and line:
            lines[int(m.group(1), 16)] = line
    if lines: break
agg = {}
tot = tots = 0.0
for r in data:
    off = int(r[iaddr], 16) - base
    key = lines.get(off, ('?', 0))
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += float(r[ia] or 0); a[1] += float(r[isamp] or 0)
    tot += float(r[ia] or 0); tots += float(r[isamp] or 0)
print(f"kernel {mangled_hint}: {tot:.0f} warp instructions" + (f", {tot*32/nsamp:.2f} thread-instr/sample" if nsamp else ""))
src = {}
for (f, l) in agg:
    if f not in src:
        for root in (os.path.dirname(os.path.abspath(so)) + '/../csrc', '.'):
            p = os.path.join(root, f)
            if os.path.exists(p):
                src[f] = open(p).read().splitlines(); break
        else: src[f] = []
print(f"{'file:line':22s} {'inst%':>6s} {'stall%':>6s}" + ("  instr/sample" if nsamp else "") + "  source")
for (f, l), (a, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ''
    print(f"{f+':'+str(l):22s} {a/tot*100:6.2f} {s/tots*100:6.2f}" + (f"  {a*32/nsamp:7.3f}    " if nsamp else "  ") + text)
