"""Scratch: per-QP statistics of the oracle core on a warm closed loop (guides the K1 design)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.models import load_model
from oracle.qp_c import CoreC
from oracle.bnb_ref import OracleController
name = sys.argv[1] if len(sys.argv) > 1 else 'cp20'
model = load_model(name)
core = CoreC(model, variant=int(os.environ.get("VARIANT", "0")))
stats = []
orig = core.solve
def solve(x0, lb, ub, warm=None):
    out = orig(x0, lb, ub, warm=warm)
    nW0 = 0 if warm is None else len(warm['rows'])
    stats.append((out['status'], out['iters'], out['prox'], nW0, len(out['warm']['rows']), int((lb == ub).sum())))
    return out
core.solve = solve
ctl = OracleController(model, core, hot_start=True)
x = np.load('warm-start-hybrid-mpc_b200/data/cp20_instances.npy')[0] if name == 'cp20' else model['x0_nominal']
rng = np.random.default_rng(0)
ws = None
for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    n0 = len(stats)
    inc, leaves, solves = ctl.feedforward(x, warm_start=ws)
    u0 = inc.primal['u'][0]
    e = 0.003 * rng.standard_normal(4) * model['x_max']
    ws = ctl.construct_warm_start(leaves, x, u0[:ctl.nuc], u0[ctl.nuc:], e)
    x = inc.primal['x'][1] + e
    s = np.array(stats[n0:])
    print('step', t, 'solves', solves, 'iters/QP %.1f' % s[:, 1].mean(), 'first-QP iters', s[0, 1], 'prox/QP %.2f' % s[:, 2].mean(),
          'nW mean %.0f max %d' % (s[:, 4].mean(), s[:, 4].max()), 'infeasible', int((s[:, 0] == 3).sum()))
    if t == 1:
        for r in s: print('   ', r)
