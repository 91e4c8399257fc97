"""Scratch: cycle breakdown of the critical path by phase (needs a -DWS_PROF build passed as WSHMPC_LIB)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model
from warm_start_hmpc_b200.closed_loop import ClosedLoop
from warm_start_hmpc_b200.capi import load_library
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
NAMES = {0: 'bnb select + bounds', 1: 'load working set', 2: 'vf0', 3: 'rebuild (rest)', 4: 'prox: Rinv matvec', 5: 'prox: bounds (rows)',
         6: 'prox: refresh_uv', 7: 'ratio test', 8: 'lam step', 10: 'pricing rows', 11: 'argmax', 13: 'dependent: ratio', 15: 'prox end: RinvT',
         16: 'y_out + pinned', 17: 'build_records', 18: 'bnb post', 19: 'shift_instance', 20: 'queue pop',
         30: 'append: row load', 31: 'append: qt_dots', 32: 'append: q_apply', 33: 'append: sums', 34: 'append: reorth', 35: 'append: ri_matvec', 36: 'append: write',
         40: 'rebuild: row load', 41: 'rebuild: qt_dots', 42: 'rebuild: q_apply', 43: 'rebuild: sums', 44: 'rebuild: reorth', 45: 'rebuild: ri_matvec', 46: 'rebuild: write',
         50: 'remove last', 51: 'remove: rotations', 52: 'remove: sweep', 53: 'remove: ri_matvec', 54: 'sweep: loads + barrier', 55: 'sweep: Ri rows', 56: 'sweep: Q1 rows + u', 60: 'prox: bounds (xi)', 62: 'pricing xi', 127: 'start'}
lib = load_library()
MODEL = os.environ.get('WS_MODEL', 'cp20')
model = load_model(MODEL)
ctl = controller_from_model(model)
from warm_start_hmpc_b200.instances import load_initial_states
if MODEL == 'cp20':
    x0 = load_initial_states(0, N)
else:
    import bench
    x0 = bench.initial_states(MODEL, model, 0, N)
rng = np.random.default_rng(1)
e = torch.as_tensor(0.003 * rng.standard_normal((6, S, N, model['A'].shape[0])) * model['x_max'], device='cuda')
L = ClosedLoop(ctl, N, warm=True, max_solves=1024 if MODEL == 'cp20' else 4096, max_roots=512 if MODEL == 'cp20' else 1024)
L.reset(x0)
buf = (C.c_ulonglong * 256)()
for w in range(6):
    if w == 4:
        torch.cuda.synchronize(); lib.wshmpc_prof_read(None, 1); b = L.totals.clone(); t0 = time.time()
    L.run(S, e=e[w])
torch.cuda.synchronize(); dt = time.time() - t0
lib.wshmpc_prof_read(buf, 0)
d = (L.totals - b).cpu().numpy()
a = np.array(buf[:], dtype=np.float64).reshape(128, 2)
tot = a[:100, 0].sum()
print('QPs %d  iters/QP %.1f  %.0f QP/s  cycles/QP (thread-0 timeline) %.0f' % (d[0], d[1] / d[0], d[0] / dt, tot / d[0]))
print('removals: mean columns rotated %.1f, mean k %.1f' % (a[100, 0] / max(a[100, 1], 1), a[101, 0] / max(a[101, 1], 1)))
a[100:102] = 0
for i in np.argsort(-a[:, 0]):
    if a[i, 0] > 0:
        print('%-24s %5.1f%%  visits/QP %6.2f  cycles/visit %8.0f' % (NAMES.get(int(i), str(i)), 100 * a[i, 0] / tot, a[i, 1] / d[0], a[i, 0] / max(a[i, 1], 1)))
