"""Scratch: fused loop with the host in the loop every step (run_mailbox) against the device-resident loop (run) and lock step.
    python tools/mailbox_timing.py [instances] [windows] [steps per window]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model, load_initial_states
from warm_start_hmpc_b200.closed_loop import ClosedLoop
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 6
S = int(sys.argv[3]) if len(sys.argv) > 3 else 20
model = load_model('cp20')
ctl = controller_from_model(model)
x0 = load_initial_states(0, N)
rng = np.random.default_rng(1)
e = 0.003 * rng.standard_normal((NW, S, N, 4)) * model['x_max']
for mode in ('device', 'mailbox'):
    L = ClosedLoop(ctl, N, warm=True, max_solves=1024, max_roots=512)
    L.reset(x0)
    for w in range(NW):
        ncall = [0]
        def plant(idx, step, u0, x1, w=w):
            ncall[0] += 1
            return x1 + e[w][step, idx], e[w][step, idx]
        torch.cuda.synchronize(); t0 = time.time(); b = L.totals.clone()
        if mode == 'device':
            logs = L.run(S, e=torch.as_tensor(e[w], device='cuda'))
        else:
            logs = L.run_mailbox(S, plant)
        torch.cuda.synchronize(); dt = time.time() - t0; d = (L.totals - b).cpu().numpy()
        print(mode, 'window', w, '%.1f ms' % (dt * 1e3), 'QPs', d[0], '%.0f QP/s' % (d[0] / dt), 'plant calls', ncall[0],
              'checksum %.12e' % float(torch.nan_to_num(logs['cost'], posinf=0.).sum()), flush=True)
