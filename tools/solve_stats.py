"""Scratch: per-solve statistics of warm-started B&B steps (needs a -DWS_PROF build passed as WSHMPC_LIB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model = load_model('cp20')
ctl = controller_from_model(model)
x = np.load('warm-start-hybrid-mpc_b200/data/cp20_instances.npy')[:N]
rng = np.random.default_rng(1)
tree = None
rows = []
for t in range(STEPS):
    n_roots = None if tree is None else tree.n_nodes.cpu().numpy().copy()
    res, tree = ctl.feedforward_batch(x, warm_start=tree, trace=True, max_solves=1024)
    torch.cuda.synchronize()
    tr = res['trace'].cpu().numpy().reshape(N, -1, 2); ns = res['n_solves'].cpu().numpy()
    for i in range(N):
        for s in range(ns[i]):
            a, b = int(tr[i, s, 0]), int(tr[i, s, 1])
            node, d = a & 0xffff, a >> 16
            rows.append((t, i, node < (1 if n_roots is None else n_roots[i]), d, b & 0xfff, (b >> 12) & 0xff, (b >> 20) & 0xff, (b >> 28) & 1, (b >> 29) & 3))
    e0 = 0.003 * rng.standard_normal((N, 4)) * model['x_max']
    tree, xn, u0 = ctl.construct_warm_start_batch(res, tree, e0=e0, max_solves=1024)
    x = xn
R = np.array(rows)
def show(name, M):
    if len(M) == 0: return
    print('%-28s n %6d  depth %5.1f  iters %6.1f (med %3d max %4d)  k_start %5.1f  k_final %5.1f  prox %.2f' % (
        name, len(M), M[:, 3].mean(), M[:, 4].mean(), np.median(M[:, 4]), M[:, 4].max(), M[:, 5].mean(), M[:, 6].mean(), M[:, 8].mean()))
W = R[R[:, 0] > 0]
show('cold step, feasible', R[(R[:, 0] == 0) & (R[:, 7] == 0)]); show('cold step, infeasible', R[(R[:, 0] == 0) & (R[:, 7] == 1)])
for root in (1, 0):
    for inf in (0, 1):
        show('warm %s %s' % ('root' if root else 'child', 'infeasible' if inf else 'feasible'), W[(W[:, 2] == root) & (W[:, 7] == inf)])
print('warm: QPs/step/instance %.2f' % (len(W) / (STEPS - 1) / N))
its = W[:, 4]
print('iteration histogram (warm):', np.histogram(its, bins=[0, 5, 10, 20, 40, 80, 160, 320, 5000])[0])
