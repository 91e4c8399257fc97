"""Scratch timing of K1 on random nodes (not a bench value)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from oracle.models import load_model
from tests.util import make_controller, random_nodes
for name, N, S in (('cp20', 1184, 296), ('cp20', 2368, 592), ('cp40', 296, 296), ('syn30', 296, 296)):
    model = load_model(name)
    ctl = make_controller(model)
    x0, lb, ub = random_nodes(model, N, seed=0)
    h = ctl.handle(n_slots=S)
    x0, lb, ub = [torch.as_tensor(a, device='cuda') for a in (x0, lb, ub)]
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        out = h.solve_nodes(x0, lb, ub)
        torch.cuda.synchronize(); dt = time.time() - t
    st = out['status'].cpu().numpy(); it = out['iters'].cpu().numpy()
    print(name, 'N', N, 'slots', S, 'smem', h.lib and None, 'time %.1f ms' % (dt * 1e3), '%.0f QP/s' % (N / dt),
          'status', {int(s): int((st == s).sum()) for s in np.unique(st)}, 'iters mean %.1f max %d' % (it.mean(), it.max()), flush=True)
