"""Scratch: a few fused closed-loop launches on a small batch, for ncu (not a bench value)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model
from warm_start_hmpc_b200.closed_loop import ClosedLoop
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = load_model('cp20')
ctl = controller_from_model(model)
x0 = np.load('warm-start-hybrid-mpc_b200/data/cp20_instances.npy')[:N]
rng = np.random.default_rng(1)
e = torch.as_tensor(0.003 * rng.standard_normal((4, S, N, 4)) * model['x_max'], device='cuda')
L = ClosedLoop(ctl, N, warm=True, max_solves=512, max_roots=256)
L.reset(x0)
for w in range(3):
    logs = L.run(S, e=e[w])
    torch.cuda.synchronize()
    print('window', w, 'QPs', int(logs['n_solves'].sum()), 'iters so far', int(L.totals[1]))
