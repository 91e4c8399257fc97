"""Scratch: nominal closed loop warm/cold on 1 instance, print status per step; hot chain over golden nodes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.models import load_model, GOLDEN
from tests.util import make_controller
from warm_start_hmpc_b200.closed_loop import ClosedLoop
model = load_model('cp20')
ctl = make_controller(model)
g = np.load(os.path.join(GOLDEN, 'cp20_nodes.npz'))
N = len(g['status'])
x0 = np.repeat(g['x0'][None], N, 0)
h = ctl.handle(n_slots=4)
for rep in range(3):
    out = h.solve_nodes(x0, g['lb'], g['ub'], slot=np.zeros(N, np.int32), hot=np.r_[0, np.ones(N - 1, np.int32)].astype(np.int32))
    st = out['status'].cpu().numpy(); cost = out['cost'].cpu().numpy(); it = out['iters'].cpu().numpy()
    bad = np.nonzero(st != g['status'])[0]
    ok = (st == 2) & (g['status'] == 2)
    print('hot chain rep', rep, 'status mismatches', bad, st[bad], 'max rel cost err', np.max(np.abs(cost[ok] - g['cost'][ok]) / np.abs(g['cost'][ok])), 'iters mean', it.mean())
gl = np.load(os.path.join(GOLDEN, 'cp20_closed_loop.npz'))
for w in (True, False):
    L = ClosedLoop(ctl, 1, warm=w, max_solves=1024, max_roots=512, n_slots=1)
    L.reset(model['x0_nominal'][None])
    for t in range(len(gl['nom_cost'])):
        out = L.step()
        print('warm' if w else 'cold', t, int(out['status'][0]), int(out['n_solves'][0]), float(out['cost'][0]), gl['nom_cost'][t])
        if int(out['status'][0]) != 0: break
