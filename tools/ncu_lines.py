"""Per-source-line aggregation of an .ncu-rep source page: samples, warp instructions, dominant stalls.
    python tools/ncu_lines.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
agg = {}; fname = '?'; hdr = None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] in ('File Name', 'File Path'):
        fname = r[1].split('/')[-1]; continue
    if len(r) > 6 and r[0] == 'Line No':
        hdr = r; ix = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) <= 6 or not r[0].strip():
        continue
    def g(name):
        try: return float(r[ix[name]])
        except Exception: return 0.
    key = (fname, int(r[0]))
    a = agg.setdefault(key, dict(src=r[1].strip()[:90], samp=0., inst=0., st={}))
    a['samp'] += g('# Samples'); a['inst'] += g('Instructions Executed')
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h:
            a['st'][h[6:]] = a['st'].get(h[6:], 0.) + g(h)
ts = sum(a['samp'] for a in agg.values()) or 1.; ti = sum(a['inst'] for a in agg.values()) or 1.
print('total samples %.0f  total warp instructions %.3e' % (ts, ti))
for (f, l), a in sorted(agg.items(), key=lambda t: -t[1]['samp'])[:top]:
    st = sorted(a['st'].items(), key=lambda t: -t[1])[:2]
    print('%5.1f%% samp %5.1f%% inst  %-16s %-28s %s' % (100 * a['samp'] / ts, 100 * a['inst'] / ti, '%s:%d' % (f[:11], l),
          ' '.join('%s=%.0f%%' % (k, 100 * v / max(a['samp'], 1)) for k, v in st), a['src']))
