"""Scratch: one K1 launch on random cp20 nodes, for ncu (not a bench value)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.models import load_model
from tests.util import make_controller, random_nodes
name = sys.argv[1] if len(sys.argv) > 1 else 'cp20'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 592
model = load_model(name)
ctl = make_controller(model)
x0, lb, ub = random_nodes(model, N, seed=0)
h = ctl.handle(n_slots=148)
x0, lb, ub = [torch.as_tensor(a, device='cuda') for a in (x0, lb, ub)]
for rep in range(2):
    out = h.solve_nodes(x0, lb, ub)
    torch.cuda.synchronize()
print('iters', out['iters'].cpu().numpy().mean())
