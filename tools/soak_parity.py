"""Scratch: larger randomized K1-vs-oracle check than the test-suite runs (status + cost), and a 100-step closed-loop soak."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.models import load_model
from oracle.qp_c import CoreC
from tests.util import make_controller, random_nodes
for name, N, seeds in (('cp20', 200, (11, 12)), ('cp40', 40, (13,)), ('syn30', 24, (14,))):
    model = load_model(name)
    ctl = make_controller(model)
    oracle = CoreC(model)
    bad = 0; worst = 0.; ninf = 0; tot = 0
    for seed in seeds:
        x0, lb, ub = random_nodes(model, N, seed=seed)
        out = ctl.handle(n_slots=64).solve_nodes(x0, lb, ub)
        st = out['status'].cpu().numpy(); cost = out['cost'].cpu().numpy()
        for i in range(N):
            ref = oracle.solve(x0[i], lb[i], ub[i])
            tot += 1
            if st[i] != ref['status']:
                bad += 1; print(name, seed, i, 'status', st[i], ref['status'])
            elif st[i] == 2:
                worst = max(worst, abs(cost[i] - ref['cost']) / abs(ref['cost']))
            else:
                ninf += 1
    print(name, 'nodes', tot, 'status mismatches', bad, 'infeasible', ninf, 'worst relative cost error %.2e' % worst, flush=True)
from warm_start_hmpc_b200.instances import controller_from_model
from warm_start_hmpc_b200.closed_loop import ClosedLoop
model = load_model('cp20')
ctl = controller_from_model(model)
N, S = 512, 100
x0 = np.load('warm-start-hybrid-mpc_b200/data/cp20_instances.npy')[:N]
rng = np.random.default_rng(3)
e = torch.as_tensor(0.003 * rng.standard_normal((S, N, 4)) * model['x_max'], device='cuda')
L = ClosedLoop(ctl, N, warm=True, max_solves=1024, max_roots=512)
L.reset(x0)
t0 = time.time(); logs = L.run(S, e=e); torch.cuda.synchronize(); dt = time.time() - t0
st = logs['status'].cpu().numpy()
print('soak: 512 instances x 100 steps in %.2f s, QPs %d (%.0f QP/s), status counts' % (dt, int(L.totals[0]), int(L.totals[0]) / dt),
      {int(k): int((st == k).sum()) for k in np.unique(st)}, 'max QPs in one step', int(logs['n_solves'].max()), 'active at the end', int(L.active.sum()))
