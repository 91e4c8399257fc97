"""Summarises an .ncu-rep (read on the CPU box) into a small text file for profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt
"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum ', 'lts__t_bytes.sum.per_second',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum ', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64', 'sm__pipe_fp64_cycles_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct', 'launch__registers_per_thread ', 'launch__block_size',
        'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit', 'smsp__cycles_active.avg ',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'smsp__inst_executed_op_shared', 'sm__cycles_elapsed.avg ']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines = []
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    lines.append('== kernel %s  grid %s block %s' % (d.get('Kernel Name', '?')[:80], d.get('Grid Size'), d.get('Block Size')))
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w.strip()) if w.endswith(' ') else (w in h) for w in WANT):
            lines.append('  %-70s %-12s %s' % (h, u, v))
    stalls = [(h, float(v)) for h, v in zip(hdr, r) if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued') and v not in ('', 'n/a')]
    tot = sum(v for _, v in stalls) or 1.
    lines.append('  -- warp stall samples (share)')
    for h, v in sorted(stalls, key=lambda t: -t[1])[:8]:
        lines.append('  %-70s %5.1f%%' % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / tot))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
try:
    agg = {}
    fname = '?'
    hdr = None
    for r in csv.reader(io.StringIO(src)):
        if len(r) == 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]; continue
        if len(r) > 6 and r[0] == 'Line No':
            hdr = r; isamp = r.index('# Samples'); continue
        if hdr is None or len(r) <= 6 or not r[0].strip():
            continue            # SASS rows have an empty line number
        try:
            v = float(r[isamp])
        except Exception:
            continue
        key = '%s:%s  %s' % (fname, r[0], r[1].strip()[:110])
        agg[key] = agg.get(key, 0.) + v
    tot = sum(agg.values()) or 1.
    lines.append('== hottest CUDA source lines by warp-stall samples (all kernels in the report)')
    for k, v in sorted(agg.items(), key=lambda t: -t[1])[:40]:
        lines.append('  %5.1f%%  %s' % (100 * v / tot, k))
except Exception as ex:
    lines.append('source page not parsed: %r' % ex)
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
