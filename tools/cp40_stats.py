"""Scratch: cold batched B&B of the horizon-40 cart-pole: iterations, working-set sizes, time per QP."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from warm_start_hmpc_b200.instances import load_model, controller_from_model
name = sys.argv[1] if len(sys.argv) > 1 else 'cp40'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148
m = load_model(name)
c = controller_from_model(m)
rng = np.random.default_rng(40)
xs = m['x0_nominal'][None] + rng.uniform(-1, 1, (N, 4)) * np.array([0.02, 0.01, 0.05, 0.05])
h = c.handle(min(N, c.default_slots()))
print('ks', 'smem bytes: see wshmpc_create', flush=True)
for rep in range(2):
    tree = c.new_tree(N, 1, 2048); h.tree_init_root(tree)
    totals = torch.zeros(8, dtype=torch.int64, device='cuda')
    xd = torch.as_tensor(xs, device='cuda')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = h.bnb_solve(xd, tree, max_solves=2048, totals=totals)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
t = totals.cpu().numpy(); ns = res['n_solves'].cpu().numpy(); st = res['status'].cpu().numpy()
print(name, 'instances', N, 'time %.3f s' % dt, 'QPs', t[0], 'QP/s %.0f' % (t[0] / dt), 'iters/QP %.1f' % (t[1] / t[0]), 'k mean %.1f max %d' % (t[2] / t[0], t[3]),
      'd mean %.1f' % (t[4] / t[0]), 'k0 mean %.1f' % (t[5] / t[0]), 'max QPs per instance', ns.max(), 'ms per QP (chain) %.2f' % (1e3 * dt / ns.max()),
      'status', {int(k): int((st == k).sum()) for k in np.unique(st)})
