"""Records, on a GPU box, the transcript tests/test_reference_replay_cpu.py replays through the UNMODIFIED reference
controller (which only exists in the authoring container, where there is no GPU):

    gpurun -- python tools/record_transcript.py        ->  gpurun_out/cp20_transcript.npz  (copy to tests/golden/)

For the first closed-loop steps of the golden noisy trajectory (tests/golden/cp20_closed_loop.npz `noisy_e`) it stores
  * every node QP the host loop asked kernel K1 through BoundedQP (inputs incl. the start of the dual method, outputs),
  * what the device-side search (K3) did at the same state: explored node sequence, leaves, bounds, incumbent,
  * what the device-side warm start (K2+K4) produced: cover, identifiers, bounds, dual = None pattern, records.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N_STEPS = 4


def ident_arrays(idents, nub, nb):
    depth = np.array([len(i) for i in idents], np.int32)
    bits = np.zeros((len(idents), nb), np.int8)
    for j, ident in enumerate(idents):
        for (t, i), v in ident.items():
            bits[j, t * nub + i] = int(v)
    return depth, bits


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    from warm_start_hmpc_b200.instances import load_model, controller_from_model
    model = load_model('cp20')
    ctl = controller_from_model(model)
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'cp20_closed_loop.npz'))
    e = gold['noisy_e'][:N_STEPS]
    nub, nb = ctl.mld.nub, ctl.problem.nb
    qp = ctl.qp
    calls = []
    launch = qp._launch

    def recording_launch(x0, lb, ub, y0, yc0):
        out = launch(x0, lb, ub, y0, yc0)
        calls.append(dict(x0=x0.copy(), lb=lb.copy(), ub=ub.copy(), hot=y0 is not None,
                          y0=np.zeros(ctl.problem.m) if y0 is None else y0.copy(),
                          yc0=np.zeros(ctl.problem.n) if yc0 is None else yc0.copy(), **{k: np.array(v) for k, v in out.items()}))
        return out
    qp._launch = recording_launch
    d = {}
    x = model['x0_nominal'].copy()
    ws_host, tree_dev = None, None
    for t in range(N_STEPS):
        c0 = len(calls)
        # host loop (reference control flow) through the facade
        ctl.device_search = False
        sol, leaves, n_qp, _ = ctl.feedforward(x, warm_start=ws_host, printing_period=None)
        # device search at the same state
        ctl.device_search = True
        res, tree = ctl.feedforward_batch(x[None], warm_start=tree_dev, trace=True, n_slots=1)
        torch.cuda.synchronize()
        assert int(res['status'][0]) == 0
        n_d = int(res['n_solves'][0])
        tr = res['trace'][0].cpu().numpy().reshape(-1, 2)[:n_d]
        bits = tree.bits[0].cpu().numpy().view(np.uint32); mask = tree.mask[0].cpu().numpy().view(np.uint32)
        order = [ctl._identifier(bits[j], mask[j]) for j in tr[:, 0]]
        dev_leaves = ctl.tree_to_leaves(tree, 0)
        assert n_d == n_qp == len(calls) - c0, (n_d, n_qp, len(calls) - c0)
        assert float(res['cost'][0]) == sol.objective
        d['k3_%d_x' % t] = x.copy()
        d['k3_%d_calls' % t] = np.array([c0, len(calls)])
        d['k3_%d_order_depth' % t], d['k3_%d_order_bits' % t] = ident_arrays(order, nub, nb)
        d['k3_%d_leaf_depth' % t], d['k3_%d_leaf_bits' % t] = ident_arrays([l.identifier for l in dev_leaves], nub, nb)
        d['k3_%d_leaf_lb' % t] = np.array([l.lb for l in dev_leaves])
        d['k3_%d_cost' % t] = float(res['cost'][0])
        d['k3_%d_primal' % t] = res['primal'][0].cpu().numpy()
        # warm start: host API (K2+K4 round trip) and device API
        uc0, ub0 = sol.variables['uc'][0], sol.variables['ub'][0]
        ws_host, _, _ = ctl.construct_warm_start(leaves, x, uc0, ub0, e[t])
        tree_dev, x_next, _ = ctl.construct_warm_start_batch(res, tree, e0=e[t][None])
        torch.cuda.synchronize()
        k4 = ctl.tree_to_leaves(tree_dev, 0)
        assert [sorted(a.identifier.items()) for a in k4] == [sorted(b.identifier.items()) for b in ws_host]
        assert np.array_equal(np.array([a.lb for a in k4]), np.array([b.lb for b in ws_host]))
        d['k4_%d_depth' % t], d['k4_%d_bits' % t] = ident_arrays([l.identifier for l in k4], nub, nb)
        d['k4_%d_lb' % t] = np.array([l.lb for l in k4])
        d['k4_%d_none' % t] = np.array([l.extra.dual is None for l in k4])
        d['k4_%d_dobj' % t] = np.array([0. if l.extra.dual is None else l.extra.dual.objective for l in k4])
        d['k4_%d_start_c' % t] = np.array([l.extra.active_set['c'] for l in k4])
        d['k4_%d_start_v' % t] = np.array([l.extra.active_set['v'] for l in k4])
        d['k4_%d_e' % t] = e[t]
        x = x_next[0].cpu().numpy().copy()
        print('step %d: %d QPs, cover %d, cost %.10f' % (t, n_d, len(k4), sol.objective), flush=True)
    for k in ('x0', 'lb', 'ub', 'hot', 'y0', 'yc0', 'status', 'cost', 'dobj', 'iters', 'primal', 'dual', 'yc'):
        d['qp_' + k] = np.array([c[k] for c in calls])
    d['n_steps'] = N_STEPS
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'cp20_transcript.npz'), **d)
    print('transcript: %d QPs, %.1f KB' % (len(calls), os.path.getsize(os.path.join(ROOT, 'gpurun_out', 'cp20_transcript.npz')) / 1e3))


if __name__ == '__main__':
    main()
