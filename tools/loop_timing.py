"""Scratch: timing of fused closed-loop windows on 512 instances (not a bench value)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model
from warm_start_hmpc_b200.closed_loop import ClosedLoop
N, S = 512, 10
model = load_model('cp20')
ctl = controller_from_model(model)
x0 = np.load('tests/golden/cp20_instances.npy')[:N]
rng = np.random.default_rng(1)
e = torch.as_tensor(0.003 * rng.standard_normal((6, S, N, 4)) * model['x_max'], device='cuda')
L = ClosedLoop(ctl, N, warm=True, max_solves=1024, max_roots=512)
L.reset(x0)
for w in range(6):
    torch.cuda.synchronize(); t0 = time.time(); b = L.totals.clone()
    logs = L.run(S, e=e[w])
    torch.cuda.synchronize(); dt = time.time() - t0; d = (L.totals - b).cpu().numpy()
    print('window', w, '%.1f ms' % (dt * 1e3), 'QPs', d[0], 'iters/QP %.1f' % (d[1] / d[0]), '%.0f QP/s' % (d[0] / dt), '%.2f us/iter/SM' % (dt * 1e6 * 148 / d[1]), 'active', int(L.active.sum()), 'k mean %.1f max %d' % (d[2] / d[0], int(L.totals[3])), flush=True)
