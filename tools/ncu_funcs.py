"""Per-function aggregation (by source line ranges) of an .ncu-rep source page."""
import csv, io, re, subprocess, sys, os
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def funcs(path):
    out = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r'^(?:template.*\n)?(?:__device__|__global__|static|extern|__host__)[^;]*?\b(\w+)\s*\(', l)
        if m and not l.strip().endswith(';'):
            out.append((i, m.group(1)))
    return out
tables = {}
for f in ('qp_device.cuh', 'records.cuh', 'bnb.cuh', 'wshmpc.cu'):
    tables[f] = funcs(os.path.join(root, 'warm-start-hybrid-mpc_b200', 'csrc', f))
def func_of(f, line):
    name = '?'
    for l0, n in tables.get(f, []):
        if l0 <= line: name = n
        else: break
    return name
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
    key = fname[:12] + ':' + func_of(fname, int(r[0]))
    a = agg.setdefault(key, dict(samp=0., inst=0., st={}))
    a['samp'] += g('# Samples'); a['inst'] += g('Instructions Executed')
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h:
            a['st'][h[6:]] = a['st'].get(h[6:], 0.) + g(h)
ts = sum(a['samp'] for a in agg.values()) or 1.; ti = sum(a['inst'] for a in agg.values()) or 1.
print('total samples %.0f  total warp instructions %.3e' % (ts, ti))
for k, a in sorted(agg.items(), key=lambda t: -t[1]['samp'])[:30]:
    st = sorted(a['st'].items(), key=lambda t: -t[1])[:3]
    print('%5.1f%% samp %5.1f%% inst  %-36s %s' % (100 * a['samp'] / ts, 100 * a['inst'] / ti, k, ' '.join('%s=%.0f%%' % (kk, 100 * v / max(a['samp'], 1)) for kk, v in st)))
