"""CPU suite: the N > 1 path (instance sharding + whole-job aggregation) with gloo, world_size 2."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from warm_start_hmpc_b200.closed_loop import shard, reduce_stats
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard(1001, r, w)
units, t = reduce_stats(hi - lo, 10. * (r + 1))
assert units == 1001, units
assert t == 10. * w, t
if r == 0:
    print('OK', units, t)
dist.destroy_process_group()
'''


def test_gloo_world2_shard_and_reduce(tmp_path):
    f = tmp_path / 'worker.py'
    f.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29533', str(f)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'OK 1001 20.0' in out.stdout


REBALANCE_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from warm_start_hmpc_b200.capi import Tree
from warm_start_hmpc_b200.rebalance import rebalance, plan_moves
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
N, nb, S, nx = 6, 80, 50, 4


def make(rank):
    g = torch.Generator().manual_seed(100 + rank)
    t = Tree(N, nb, S, 12, 10, 'cpu')
    t.n_nodes[:] = torch.randint(3, 9, (N,), generator=g, dtype=torch.int32)
    t.n_recs[:] = t.n_nodes
    for name in ('depth', 'alive', 'rec'):
        getattr(t, name)[:] = torch.randint(0, 50, getattr(t, name).shape, generator=g, dtype=torch.int32)
    t.bits[:] = torch.randint(0, 2 ** 30, t.bits.shape, generator=g, dtype=torch.int32)
    t.lb[:] = torch.rand(t.lb.shape, generator=g, dtype=torch.float64)
    t.rec_dobj[:] = torch.rand(t.rec_dobj.shape, generator=g, dtype=torch.float64)
    t.rec_dual[:] = torch.rand(t.rec_dual.shape, generator=g, dtype=torch.float64)
    x = torch.rand((N, nx), generator=g, dtype=torch.float64)
    active = torch.ones(N, dtype=torch.int32)
    if rank == 1:
        active[[0, 2, 3, 5]] = 0                      # rank 1 lost four instances
    gid = torch.arange(N, dtype=torch.int64) + 1000 * rank
    return t, x, active, gid


assert plan_moves([6, 2], [0, 4]) == [(0, 1, 2)]
t, x, active, gid = make(r)
sent, got, plan = rebalance(t, x, active, gid)
assert plan == [(0, 1, 2)], plan
if r == 0:
    assert (sent, got) == (2, 0) and int(active.sum()) == 4 and active[4] == 0 and active[5] == 0
else:
    assert (sent, got) == (0, 2) and int(active.sum()) == 4
    t0, x0, a0, g0 = make(0)                          # what rank 0 held: instances 4 and 5 must have arrived in slots 0 and 2
    for slot, src in ((0, 4), (2, 5)):
        nn = int(t0.n_nodes[src])
        assert int(t.n_nodes[slot]) == nn and int(t.n_recs[slot]) == nn and int(gid[slot]) == src
        assert torch.equal(x[slot], x0[src])
        for name in ('depth', 'alive', 'rec', 'bits', 'lb'):
            assert torch.equal(getattr(t, name)[slot, :nn], getattr(t0, name)[src, :nn]), name
        assert torch.equal(t.rec_dobj[slot, :nn], t0.rec_dobj[src, :nn]) and torch.equal(t.rec_dual[slot, :nn], t0.rec_dual[src, :nn])
    assert active[0] == 1 and active[2] == 1 and active[3] == 0
# a second call finds the ranks balanced
assert rebalance(t, x, active, gid) == (0, 0, [])
if r == 0:
    print('REBALANCE OK')
dist.destroy_process_group()
'''


SPLIT_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from oracle.models import load_model, GOLDEN
from tests.util import make_controller, oracle_launch
from warm_start_hmpc_b200.split_frontier import split_frontier_bnb
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
model = load_model('cp20')
ctl = make_controller(model)
pd = ctl.problem
launch = oracle_launch(model, pd)


def solve(x0, nodes):
    """the device call of GpuBatchSolver replaced by the CPU oracle core (no GPU in this suite)"""
    rows = []
    for node in nodes:
        lb, ub = ctl._get_bound_binaries(node.identifier)
        y0 = yc0 = None
        if node.extra is not None and node.extra.active_set is not None:
            y0 = ctl.qp._signed_multipliers(np.asarray(node.extra.active_set['c'])); yc0 = np.asarray(node.extra.active_set['v'])[:pd.n]
        o = launch(np.asarray(x0, float), lb.ravel(), ub.ravel(), y0, yc0)
        rows.append(np.concatenate(([o['status'], o['cost'], o['dobj']], o['dual'], o['yc'], o['primal'])))
    width = 3 + pd.layout.dual + pd.n + pd.layout.primal
    return torch.as_tensor(np.array(rows, dtype=float).reshape(len(nodes), width))


g = np.load(os.path.join(GOLDEN, 'cp20_nodes.npz'))
sol, leaves, solves, rounds = split_frontier_bnb(ctl, g['x0'], solve, nodes_per_rank=2)
assert abs(sol.objective - float(g['opt_cost'])) <= 1e-6 * float(g['opt_cost']), (sol.objective, float(g['opt_cost']))
assert np.array_equal(np.array(sol.variables['ub']), g['opt_ub'])
assert solves >= len(g['status']) - 2 and rounds < solves
# every rank holds the same frontier: compare a digest
digest = torch.tensor([float(len(leaves)), float(sum(l.lb for l in leaves if np.isfinite(l.lb))), float(solves), float(rounds)], dtype=torch.float64)
both = [torch.zeros_like(digest) for _ in range(w)]
dist.all_gather(both, digest)
assert all(torch.equal(both[0], b) for b in both)
# the leaves are a cover the warm start can shift (host formulation: no GPU here)
ctl.device_search = False
ws, _, _ = ctl.construct_warm_start(leaves, g['x0'], sol.variables['uc'][0], sol.variables['ub'][0], np.zeros(4))
assert len(ws) >= 70
if r == 0:
    print('SPLIT OK', solves, rounds, len(leaves))
dist.destroy_process_group()
'''


def _run_world2(tmp_path, body, port, timeout=600):
    f = tmp_path / 'worker.py'
    f.write_text(body % ROOT)
    return subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                           '--master-addr', '127.0.0.1', '--master-port', str(port), str(f)],
                          capture_output=True, text=True, timeout=timeout)


def test_gloo_world2_rebalance_moves_instances_with_their_trees(tmp_path):
    """Load balancing (north_star item 4): two live instances move from the full rank into dead slots of the other one,
    with x, identity and the warm-start tree (nodes + dual records) arriving bit for bit."""
    out = _run_world2(tmp_path, REBALANCE_WORKER, 29541)
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'REBALANCE OK' in out.stdout


def test_gloo_world2_split_frontier_incumbent_allreduce(tmp_path):
    """Split-frontier B&B of one MIQP over two ranks (incumbent min-allreduce + replicated frontier): same optimum and
    mode sequence as the golden sequential run, identical frontier on both ranks, leaves usable as a warm start."""
    out = _run_world2(tmp_path, SPLIT_WORKER, 29543)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    assert 'SPLIT OK' in out.stdout
