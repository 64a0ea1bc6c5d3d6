#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + instruction mix per SASS region.  usage: ncu_summary.py rep nsamples"""
import csv, subprocess, sys, io
rep=sys.argv[1]; nsamp=float(sys.argv[2]) if len(sys.argv)>2 else None
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr,units,vals=rows[0],rows[1],rows[2]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size']
for i,h in enumerate(hdr):
    if h in want: print(f'{h:75s} {vals[i]:>18s} {units[i]}')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src))); hdr=rows[1]; data=rows[2:]
ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples')
tot=sum(float(r[ia] or 0) for r in data); tots=sum(float(r[isamp] or 0) for r in data)
print('total warp inst',tot, 'thread-inst per sample', tot*32/nsamp if nsamp else '')
B=int(sys.argv[3]) if len(sys.argv)>3 else 96
for b in range(0,len(data),B):
    seg=data[b:b+B]
    v=sum(float(r[ia] or 0) for r in seg)/tot*100; s=sum(float(r[isamp] or 0) for r in seg)/tots*100
    ops={}
    for r in seg:
        t=r[isrc].split()
        if not t: continue
        op=t[0] if not t[0].startswith('@') else (t[1] if len(t)>1 else t[0])
        op=op.split('.')[0]; ops[op]=ops.get(op,0)+float(r[ia] or 0)
    top=sorted(ops.items(), key=lambda kv:-kv[1])[:6]
    if v>0.7 or s>0.7: print(f"rows {b:5d}-{b+B:5d}: inst {v:5.1f}%  stall-samples {s:5.1f}%  ", ' '.join(f"{k}:{x/tot*100:.1f}" for k,x in top))
