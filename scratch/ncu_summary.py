"""Summarise an .ncu-rep of the fused kernel: headline metrics, per-phase instruction shares
(phases delimited by barrier instructions in SASS order) and opcode histogram."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:85s} {r[i]:>16s} {units[i]}")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        try:
            if float(r[i]) > 0.15: print(f"{h:85s} {r[i]:>16s}")
        except ValueError: pass
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]
isrc, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = []
for rr in rows[2:]:
    if rr and rr[0] == 'Kernel Name': break
    try: data.append((rr[isrc].strip(), int(rr[ie] or 0), int(rr[isamp] or 0)))
    except (ValueError, IndexError): pass
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print(f"SASS instructions {len(data)}, executed warp-instr {tot}, samples {ts}")
si = ss = 0; start = 0
for i, d in enumerate(data):
    si += d[1]; ss += d[2]
    if any(m in d[0] for m in ('BAR.SYNC', 'BAR.ARV', 'EXIT')):
        if si > 0.003 * tot: print(f"  sass {start:5d}-{i:5d}  inst {100*si/tot:5.1f}%  samples {100*ss/ts:5.1f}%  ends: {d[0][:50]}")
        start = i + 1; si = ss = 0
h = collections.Counter()
for d in data:
    t = d[0].split()
    if not t: continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    h[op.split('.')[0]] += d[1]
print('  '.join(f"{op} {100*c/tot:.1f}%" for op, c in h.most_common(24)))
