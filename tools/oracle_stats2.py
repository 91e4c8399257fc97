"""Scratch: iterations per QP when each node is hot-started from ITS OWN dual record (parent / shifted leaf)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.models import load_model
from oracle.qp_c import CoreC
from oracle.bnb_ref import OracleController
mode = sys.argv[1]
nsteps = int(sys.argv[2])
model = load_model('cp20')
core = CoreC(model, variant=1)
stats = []
class Ctl(OracleController):
    def solve_node(self, node, x0):
        c = self.c
        if mode == 'record':
            if node.dual is None:
                self._warm = None
            else:
                var = node.dual['variables']
                rows, sides, lam = [], [], []
                for t in range(self.T):
                    for i in np.nonzero(var['mu'][t] > 0)[0]:
                        rows.append(c.row0[t] + i); sides.append(1); lam.append(var['mu'][t][i])
                    for i in range(self.nub):
                        if var['nu_ub'][t][i] > 0: rows.append(c.mc + t * self.nub + i); sides.append(1); lam.append(var['nu_ub'][t][i])
                        elif var['nu_lb'][t][i] > 0: rows.append(c.mc + t * self.nub + i); sides.append(-1); lam.append(var['nu_lb'][t][i])
                self._warm = dict(rows=rows, sides=sides, lam=lam, z=(node.dual.get('yc') if os.environ.get('USE_YC', '1') == '1' else None))
        elif mode == 'none':
            self._warm = None
        nW0 = 0 if self._warm is None else len(self._warm['rows'])
        self.hot_start = True
        OracleController.solve_node(self, node, x0)
        if node.dual is not None and self._warm is not None and node.primal is not None:
            node.dual['yc'] = self._warm['z']
        stats.append((node.primal is not None, self._last['iters'], self._last['prox'], nW0, len(self._warm['rows'])))
orig = core.solve
def solve(x0, lb, ub, warm=None):
    out = orig(x0, lb, ub, warm=warm)
    ctl._last = out
    return out
core.solve = solve
ctl = Ctl(model, core, hot_start=True)
tot = []
for inst in range(3):
    x = np.load('warm-start-hybrid-mpc_b200/data/cp20_instances.npy')[inst]
    rng = np.random.default_rng(inst)
    ws = None
    for t in range(nsteps):
        n0 = len(stats)
        inc, leaves, solves = ctl.feedforward(x, warm_start=ws)
        if inc is None: break
        u0 = inc.primal['u'][0]
        e = 0.003 * rng.standard_normal(4) * model['x_max']
        ws = ctl.construct_warm_start(leaves, x, u0[:ctl.nuc], u0[ctl.nuc:], e)
        x = inc.primal['x'][1] + e
        s = np.array(stats[n0:], dtype=float)
        print(inst, 'step', t, 'solves', solves, 'iters/QP %.1f' % s[:, 1].mean(), 'nW0 %.0f' % s[:, 3].mean(), 'total work (iters + rebuild) %.0f' % (s[:, 1].sum() + s[:, 3].sum()),
              'cost %.9f' % inc.primal['objective'])
