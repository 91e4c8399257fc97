"""Scratch: timing of fused closed-loop windows (not a bench value).
    python tools/loop_timing.py [instances] [windows] [steps per window]        (WSHMPC_LIB / WSHMPC_LANES select the build / lanes)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model, load_initial_states
from warm_start_hmpc_b200.closed_loop import ClosedLoop
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 6
S = int(sys.argv[3]) if len(sys.argv) > 3 else 10
MODEL = os.environ.get('WS_MODEL', 'cp20')
model = load_model(MODEL)
ctl = controller_from_model(model)
if MODEL == 'cp20':
    x0 = load_initial_states(0, N)
else:
    import bench
    x0 = bench.initial_states(MODEL, model, 0, N)
rng = np.random.default_rng(1)
e = torch.as_tensor(0.003 * rng.standard_normal((NW, S, N, model['A'].shape[0])) * model['x_max'], device='cuda')
L = ClosedLoop(ctl, N, warm=True, max_solves=1024 if MODEL == 'cp20' else 2048, max_roots=512 if MODEL == 'cp20' else 1024)
print('lanes per CTA', ctl.default_slots() // 148, 'slots', L.h.n_slots, flush=True)
L.reset(x0)
for w in range(NW):
    torch.cuda.synchronize(); t0 = time.time(); b = L.totals.clone()
    logs = L.run(S, e=e[w])
    torch.cuda.synchronize(); dt = time.time() - t0; d = (L.totals - b).cpu().numpy()
    print('window', w, '%.1f ms' % (dt * 1e3), 'QPs', d[0], 'iters/QP %.1f' % (d[1] / d[0]), '%.0f QP/s' % (d[0] / dt), '%.2f us/iter/SM' % (dt * 1e6 * 148 / d[1]), 'active', int(L.active.sum()), 'k mean %.1f max %d' % (d[2] / d[0], int(L.totals[3])), 'checksum %.12e' % float(torch.nan_to_num(logs['cost'], posinf=0.).sum()), flush=True)
