"""TEST INFRASTRUCTURE (oracle side) -- golden node QPs of the synthetic MLD system (BASELINE configs[4]).

Random identifiers of SYN30 are almost always infeasible, so the feasible side of K1 on the large synthetic
system is pinned with nodes of a DIVE: starting at the root, the next binary (chronological order,
controller.py:13-44) is pinned to the value whose child is feasible with the lower cost; both children of
every level are stored with the oracle's status and cost (oracle/qp_core.c, certificates checked in
tests/test_oracle_cpu.py).  The reference has no generator for this system, so there is no
reference output to compare with; parity here is CUDA vs oracle on identical bits.

    python -m oracle.make_syn30_nodes        -> tests/golden/syn30_nodes.npz
"""
import os
import numpy as np
from oracle.models import load_model, GOLDEN
from oracle.qp_c import CoreC


def main(levels=14):
    model = load_model('syn30')
    core = CoreC(model)
    T, nub = int(model['T']), int(model['nub'])
    nb = T * nub
    x0 = model['x0_nominal']
    lb = np.zeros(nb); ub = np.ones(nb)
    rows = []
    r = core.solve(x0, lb, ub)
    rows.append((lb.copy(), ub.copy(), r['status'], r['cost'] if r['status'] == 2 else np.inf))
    for d in range(levels):
        kids = []
        for v in (0., 1.):
            l2, u2 = lb.copy(), ub.copy(); l2[d] = u2[d] = v
            r = core.solve(x0, l2, u2)
            kids.append((l2, u2, r['status'], r['cost'] if r['status'] == 2 else np.inf))
        rows.extend(kids)
        feas = [k for k in kids if k[2] == 2]
        if not feas:
            break
        best = min(feas, key=lambda k: k[3])
        lb, ub = best[0].copy(), best[1].copy()
    np.savez_compressed(os.path.join(GOLDEN, 'syn30_nodes.npz'), x0=x0, lb=np.array([k[0] for k in rows]),
                        ub=np.array([k[1] for k in rows]), status=np.array([k[2] for k in rows], dtype=np.int32),
                        cost=np.array([k[3] for k in rows]))
    print('%d nodes, %d feasible' % (len(rows), sum(k[2] == 2 for k in rows)))


if __name__ == '__main__':
    main()
