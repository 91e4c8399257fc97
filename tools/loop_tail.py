"""Scratch: is a fused window bound by the longest instance chain? (not a bench value)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model
from warm_start_hmpc_b200.closed_loop import ClosedLoop
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
W = int(sys.argv[3]) if len(sys.argv) > 3 else 6
model = load_model('cp20')
ctl = controller_from_model(model)
x0 = np.load('warm-start-hybrid-mpc_b200/data/cp20_instances.npy')
x0 = x0[np.arange(N) % len(x0)]
rng = np.random.default_rng(1)
e = torch.as_tensor(0.003 * rng.standard_normal((W, S, N, 4)) * model['x_max'], device='cuda')
L = ClosedLoop(ctl, N, warm=True, max_solves=1024, max_roots=512)
L.reset(x0)
print('slots', L.h.n_slots)
for w in range(W):
    torch.cuda.synchronize(); t0 = time.time(); b = L.totals.clone()
    logs = L.run(S, e=e[w])
    torch.cuda.synchronize(); dt = time.time() - t0; d = (L.totals - b).cpu().numpy()
    ns = logs['n_solves'].cpu().numpy().sum(axis=0)
    print('window', w, '%.1f ms' % (dt * 1e3), 'QPs', d[0], '%.0f QP/s' % (d[0] / dt), 'per-instance QPs mean %.1f max %d' % (ns.mean(), ns.max()),
          'ms per QP if chain-bound %.3f, if throughput-bound %.3f' % (dt * 1e3 / ns.max(), dt * 1e3 * min(L.h.n_slots, N) / d[0]), flush=True)
