"""Scratch: K2+K4 (shift_tree_kernel) alone on the trees a warm-started loop leaves behind -- bench.py's roofline_k2k4 leg.
    python tools/k2k4_timing.py [cp20|cp40] [instances]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from warm_start_hmpc_b200.instances import load_model, controller_from_model
from warm_start_hmpc_b200.closed_loop import ClosedLoop
W = sys.argv[1] if len(sys.argv) > 1 else 'cp20'
cfg = bench.WORKLOADS[W]
N = int(sys.argv[2]) if len(sys.argv) > 2 else cfg['instances']
model = load_model(cfg['model']); ctl = controller_from_model(model)
x0 = bench.initial_states(W, model, 0, N)
S = cfg['window']
e = torch.as_tensor(bench.noise(model, 3 * S, N, 0.003, 1), device='cuda').reshape(3, S, N, -1)
L = ClosedLoop(ctl, N, warm=True, max_solves=cfg['max_solves'], max_roots=cfg['max_roots'])
L.reset(x0)
for w in range(3):
    L.run(S, e=e[w])
torch.cuda.synchronize()
r = bench.k2k4_roofline(torch, L, ctl.problem, 6548.8)
print(W, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != 'note'}, 'checksum %.12e' % float(L.trees[1 - L.cur].lb.nan_to_num(posinf=0., neginf=0.).sum()))
