"""Scratch: find the node on which the hot-started device B&B disagrees with a from-scratch solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.models import load_model
from oracle.qp_c import CoreC
from tests.util import make_controller
from warm_start_hmpc_b200.closed_loop import ClosedLoop
model = load_model('cp20')
ctl = make_controller(model)
L = ClosedLoop(ctl, 1, warm=False, max_solves=1024, max_roots=512, n_slots=1)
L.reset(model['x0_nominal'][None])
L.step(); L.step()
x = L.x.cpu().numpy().copy()
print('state', repr(x))
res, tree = ctl.feedforward_batch(x, trace=True, n_slots=1)
n = int(res['n_solves'][0]); print('status', int(res['status'][0]), 'solves', n)
tr = res['trace'][0].cpu().numpy().reshape(-1, 2)[:n]
depth = tree.depth[0].cpu().numpy(); bits = tree.bits[0].cpu().numpy().view(np.uint32); lbt = tree.lb[0].cpu().numpy()
nb = ctl.problem.nb
lb = np.zeros((n, nb)); ub = np.ones((n, nb))
for q, (j, it) in enumerate(tr):
    d = depth[j]
    v = np.array([(bits[j, b >> 5] >> (b & 31)) & 1 for b in range(d)], dtype=float)
    lb[q, :d] = v; ub[q, :d] = v
h = ctl.handle(n_slots=148)
out = h.solve_nodes(np.repeat(x, n, 0), lb, ub)
st = out['status'].cpu().numpy(); cost = out['cost'].cpu().numpy(); its = out['iters'].cpu().numpy()
oracle = CoreC(model, variant=1)
for q, (j, it) in enumerate(tr):
    hot_cost = lbt[j]     # lb of the node after its solve (children bounds are elsewhere)
    flag = ''
    if np.isinf(hot_cost) != np.isinf(cost[q]) or (np.isfinite(cost[q]) and abs(hot_cost - cost[q]) > 1e-6 * abs(cost[q])):
        ref = oracle.solve(x[0], lb[q], ub[q])
        flag = '  <-- MISMATCH oracle status %d cost %r' % (ref['status'], ref.get('cost'))
    if flag or q > n - 4:
        print(q, 'node', j, 'depth', depth[j], 'hot iters', it, 'hot lb', hot_cost, '| scratch status', st[q], 'cost', cost[q], 'iters', its[q], flag)
