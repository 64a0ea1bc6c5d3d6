#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: key raw metrics, dram bytes per sample, opcode mix, top source lines.
usage: profile_summary.py report.ncu-rep lib.so kernel_mangled_substring n_samples [launch_name_filter]"""
import csv, io, subprocess, sys, collections
rep, so, kname, nsamp = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
pick = sys.argv[5] if len(sys.argv) > 5 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
vals = [r for r in rows[2:] if pick is None or pick in r[hdr.index('Kernel Name')]][-1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
print(f"# ncu --set full --clock-control none, one launch; {nsamp:.0f} input samples in this launch")
g = {}
for k in want:
    if k in hdr:
        i = hdr.index(k); g[k] = vals[i]
        print(f"{k:72s} {vals[i]:>20s} {units[i]}")
def num(k):
    return float(g[k].replace(',', ''))
def scale(k):
    u = units[hdr.index(k)].lower()
    return {'gbyte': 1e9, 'mbyte': 1e6, 'kbyte': 1e3, 'byte': 1.0, 'tbyte': 1e12}.get(u, 1.0)
db = num('dram__bytes_read.sum') * scale('dram__bytes_read.sum') + num('dram__bytes_write.sum') * scale('dram__bytes_write.sum')
print(f"dram_bytes_per_sample {db / nsamp:.4f}   # dram__bytes_read.sum + dram__bytes_write.sum per input sample (algorithmic: sizeof(sample) + output)")
print(f"dram_bytes_per_launch {db:.0f}")
print(f"thread_instructions_per_sample {num('smsp__inst_executed.sum') * 32 / nsamp:.2f}")
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
cand = [i for i in starts if pick is None or pick in rows[i][1]]
i0 = cand[-1]; i1 = min([j for j in starts if j > i0] + [len(rows)])
h2 = rows[i0 + 1]; data = [r for r in rows[i0 + 2:i1] if r]
ia, isrc = h2.index('Instructions Executed'), h2.index('Source')
ops = collections.Counter()
for r in data:
    t = r[isrc].split(); op = t[0] if not t[0].startswith('@') else t[1]
    ops[op.split('.')[0]] += float(r[ia] or 0)
print("# opcode mix, thread-instructions per sample")
print('  '.join(f"{k}:{v * 32 / nsamp:.2f}" for k, v in ops.most_common(14)))
stall = collections.Counter()
for i, hname in enumerate(h2):
    if hname.startswith('stall_') and 'Not Issued' not in hname:
        stall[hname[6:]] = sum(float(r[i] or 0) for r in data)
tot = sum(stall.values()) or 1
print("# warp stall samples, share")
print('  '.join(f"{k}:{v / tot * 100:.1f}%" for k, v in stall.most_common(8)))
print("# top source lines")
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace('profile_summary.py', 'ncu_lines.py'), rep, so, kname, str(nsamp), '14'] + ([pick] if pick else []))
