"""Summarise an .ncu-rep (raw page + per-opcode warp-state samples from the source page) as text.

    python tools/ncu_summary.py gpurun_out/step_full_r1b.ncu-rep > profiles/....txt
"""
import csv, io, subprocess, sys, collections

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum']

def page(rep, name):
    return subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout

rep = sys.argv[1]
rows = list(csv.reader(io.StringIO(page(rep, 'raw'))))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('Kernel Name = ' + r[hdr.index('Kernel Name')])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'  {k} [{units[i]}] = {r[i]}')
src = page(rep, 'source')
lines = list(csv.reader(io.StringIO(src)))
hd, by, tot, kname = None, None, 0.0, ''
def flush():
    if by and tot > 0:
        print(f'# warp-state samples by opcode (source page) for: {kname}; total {tot:.0f}')
        for o, (s, e) in sorted(by.items(), key=lambda kv: -kv[1][0])[:12]:
            print(f'   {o:10s} {100 * s / tot:5.1f}%  exec={e:.4g}')
for l in lines:
    if len(l) >= 2 and l[0] == 'Kernel Name':
        flush(); kname = l[1]; hd = None; by = collections.defaultdict(lambda: [0.0, 0.0]); tot = 0.0
        continue
    if '# Samples' in l and 'Source' in l:
        flush(); hd = l; ix = {k: i for i, k in enumerate(hd)}; by = collections.defaultdict(lambda: [0.0, 0.0]); tot = 0.0
        continue
    if hd is None or len(l) != len(hd):
        continue
    try:
        s = float(l[ix['# Samples']]); e = float(l[ix['Instructions Executed']])
    except ValueError:
        continue
    op = [t for t in l[ix['Source']].strip().split() if not t.startswith('@')]
    if not op:
        continue
    o = op[0].split('.')[0]
    by[o][0] += s; by[o][1] += e; tot += s
flush()
