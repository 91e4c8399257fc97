"""Launch list of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log: per kernel
the launches, total time, share of the GPU time and DRAM bytes; per launch of the dominant kernel its time and DRAM traffic.
    python tools/launch_list.py gpurun_out/x_launches.csv profiles/x_launch_list.txt [profiles/closed_loop_kernel_traffic.json]
Times under ncu are serialised and cold-cache: compare SHARES, not absolutes."""
import csv, json, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; I = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[I['ID']]), {'name': r[I['Kernel Name']], 'grid': r[I['Grid Size']], 'block': r[I['Block Size']]})
    d[r[I['Metric Name']]] = float(r[I['Metric Value']].replace(',', ''))
per = {}
for i in sorted(launches):
    d = launches[i]
    k = per.setdefault(d['name'].split('(')[0][:60], [0, 0., 0., 0.])
    k[0] += 1; k[1] += d.get('gpu__time_duration.sum', 0.); k[2] += d.get('dram__bytes_read.sum', 0.); k[3] += d.get('dram__bytes_write.sum', 0.)
tot = sum(v[1] for v in per.values()) or 1.
out = ['%s' % ' '.join(sys.argv[1:2]), '%-62s %8s %12s %8s %12s %12s' % ('kernel', 'launches', 'time ms', 'share', 'DRAM rd GB', 'DRAM wr GB')]
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
    out.append('%-62s %8d %12.3f %7.2f%% %12.3f %12.3f' % (k, v[0], v[1] / 1e6, 100 * v[1] / tot, v[2] / 1e9, v[3] / 1e9))
top = max(per, key=lambda k: per[k][1])
out.append('')
out.append('launches of %s:' % top)
big = []
for i in sorted(launches):
    d = launches[i]
    if d['name'].startswith(top):
        out.append('  id %3d  grid %-14s block %-14s %10.3f ms  DRAM read %8.3f GB  write %8.3f GB' % (
            i, d['grid'], d['block'], d['gpu__time_duration.sum'] / 1e6, d.get('dram__bytes_read.sum', 0) / 1e9, d.get('dram__bytes_write.sum', 0) / 1e9))
        big.append(d)
open(sys.argv[2], 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
if len(sys.argv) > 3:
    tag = sys.argv[4] if len(sys.argv) > 4 else ''
    timed = big[int(sys.argv[5])] if len(sys.argv) > 5 else big[-1]
    json.dump({'kernel': top, 'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on the bench command itself (%s): the %s launch of the '
               'kernel = one bench step (512 instances x 20 MPC steps, steady state)' % (sys.argv[1], tag),
               'dram_bytes_read': timed['dram__bytes_read.sum'], 'dram_bytes_write': timed['dram__bytes_write.sum'],
               'dram_bytes_per_launch': timed['dram__bytes_read.sum'] + timed['dram__bytes_write.sum'],
               'ms_under_ncu': timed['gpu__time_duration.sum'] / 1e6}, open(sys.argv[3], 'w'), indent=1)
