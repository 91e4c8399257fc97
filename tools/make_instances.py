"""Generates warm-start-hybrid-mpc_b200/data/cp20_instances.npy on a GPU box: candidates uniform in
+-[0.35, 0.2, 1.0, 0.6] from np.random.default_rng(0), the first 4096 whose step-0 MIQP is feasible.
    gpurun -- python tools/make_instances.py    ->  gpurun_out/cp20_instances.npy
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from warm_start_hmpc_b200.instances import load_model, controller_from_model
model = load_model('cp20')
ctl = controller_from_model(model)
rng = np.random.default_rng(0)
cand = rng.uniform(-1, 1, (16384, 4)) * np.array([0.35, 0.2, 1.0, 0.6])
keep = []
for lo in range(0, len(cand), 2048):
    res, tree = ctl.feedforward_batch(cand[lo:lo + 2048], max_solves=2048)
    st = res['status'].cpu().numpy(); ns = res['n_solves'].cpu().numpy()
    print(lo, {int(k): int((st == k).sum()) for k in np.unique(st)}, 'mean solves', ns.mean(), flush=True)
    keep.append(cand[lo:lo + 2048][st == 0])
    del res, tree
    if sum(len(k) for k in keep) >= 4096:
        break
x = np.concatenate(keep)[:4096]
os.makedirs('gpurun_out', exist_ok=True)
np.save('gpurun_out/cp20_instances.npy', x)
print('saved', x.shape)
