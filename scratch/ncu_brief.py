"""Key counters of every kernel in an .ncu-rep (ncu --set full), one block per kernel: duration, DRAM bytes, issue /
tensor / data-pipe utilisation, shared-memory wavefronts split into LSU (LDS / STS / TMEM loads) and tensor-core operand
reads, top stall reasons.   python scratch/ncu_brief.py report.ncu-rep > profiles/summary.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name) if name in hdr else None
want = [('gpu__time_duration.sum', 'duration'), ('sm__cycles_elapsed.avg', 'SM cycles elapsed'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers/thread'),
        ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('smsp__inst_executed.sum', 'warp instructions'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active % (ncu)'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'data pipe: LSU shared wavefronts (LDS+STS+LDTM)'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', '   of which loads (LDS, LDTM)'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', '   of which stores'),
        ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'data pipe: tensor-core operand wavefronts'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'data pipe LSU share %'),
        ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'data pipe tensor share %'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', '"bank conflicts" ld (counts TMEM-load wavefronts)'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'bank conflicts st')]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    print('=' * 100)
    print(r[col('Kernel Name')])
    for name, label in want:
        i = col(name)
        if i is not None: print(f"  {label:58s} {r[i]:>18s} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            try: st.append((float(r[i]), h.split('stalled_')[1].split('_per_')[0]))
            except ValueError: pass
    print("  stall reasons (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:6]))
