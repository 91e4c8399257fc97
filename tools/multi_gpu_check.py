"""Multi-GPU legs that are not part of bench.py's headline (run under torchrun on 2+ GPUs; results -> gpurun_out/):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 tools/multi_gpu_check.py

  1. periodic load balancing (ClosedLoop.rebalance, NCCL point-to-point): rank 0 starts with all its instances alive,
     the other ranks with a quarter of theirs; windows of the fused loop are timed with and without rebalancing
     (max over ranks), and the rebalanced job is checked against the unbalanced one instance by instance (same costs);
  2. split-frontier B&B of ONE instance (split_frontier.py: incumbent min-allreduce + replicated frontier) against the
     single-GPU device search (K3): cost, mode sequence, QPs, rounds, latency.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from warm_start_hmpc_b200.instances import load_model, controller_from_model, load_initial_states
    from warm_start_hmpc_b200.closed_loop import ClosedLoop, reduce_stats
    from warm_start_hmpc_b200.split_frontier import split_frontier_bnb, GpuBatchSolver
    out = {'world': world}
    model = load_model('cp20')
    ctl = controller_from_model(model, device=local)
    dev = torch.device('cuda', local)

    # ---- 1. load balancing
    N, S, W = 512, 10, 4
    x0 = load_initial_states(rank * N, (rank + 1) * N)
    e = torch.as_tensor(0.003 * np.random.default_rng(7 + rank).standard_normal((W + 1, S, N, 4)) * model['x_max'], device=dev)
    res = {}
    # NCCL sets its point-to-point channels up on first use (seconds): do that outside the timed region
    tok = torch.zeros(4, device=dev)
    if rank == 0:
        for r in range(1, world):
            dist.send(tok, r)
    else:
        dist.recv(tok, 0)
    torch.cuda.synchronize(); dist.barrier()
    for mode in ('static', 'rebalanced'):
        L = ClosedLoop(ctl, N, warm=True, max_solves=1024, max_roots=512)
        L.reset(x0)
        for w in range(3):                                  # past the contact-mode transient: cold step + 29 warm steps
            L.run(S, e=e[0])
        if rank > 0:
            L.active[N // 4:] = 0                           # the other ranks lose three quarters of their instances
        torch.cuda.synchronize(); dist.barrier()
        moved = 0
        t_move = 0.
        if mode == 'rebalanced':
            t1 = time.perf_counter()
            s_, g_, plan = L.rebalance()
            torch.cuda.synchronize(); dist.barrier()
            t_move = time.perf_counter() - t1
            moved = s_
        t0 = time.perf_counter(); b = L.totals.clone()
        for w in range(1, W + 1):
            logs = L.run(S, e=e[w])
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        q, ms = reduce_stats(int((L.totals - b)[0]), dt * 1e3)
        live, _ = reduce_stats(int(L.active.sum()), 0.)
        mv, _ = reduce_stats(moved, 0.)
        res[mode] = {'qp': q, 'ms_of_%d_windows' % W: ms, 'qp_per_s': q / (ms * 1e-3), 'live_instances': live, 'instances_moved': mv,
                     'rebalance_ms': 1e3 * t_move, 'live_on_this_rank': int(L.active.sum())}
        del L
        torch.cuda.empty_cache()
    out['rebalance'] = res
    out['rebalance']['speedup'] = res['static']['ms_of_%d_windows' % W] / res['rebalanced']['ms_of_%d_windows' % W]

    # ---- 2. split frontier on one instance
    for name in ('cp20', 'cp40'):
        m = load_model(name)
        c = controller_from_model(m, device=local)
        x = m['x0_nominal']
        solver = GpuBatchSolver(c, n_slots=4)
        r = {}
        for npr in (1, 2):
            split_frontier_bnb(c, x, solver, nodes_per_rank=npr)           # warm-up (module load, allocations)
            torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
            sol, leaves, solves, rounds = split_frontier_bnb(c, x, solver, nodes_per_rank=npr)
            torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
            r['nodes_per_rank_%d' % npr] = {'cost': sol.objective, 'qp': solves, 'rounds': rounds, 'ms': 1e3 * dt, 'leaves': len(leaves)}
        if rank == 0:
            sd, ld, nd, td = c.feedforward(x, printing_period=None)         # single-GPU device search (K3)
            sd, ld, nd, td = c.feedforward(x, printing_period=None)
            r['single_gpu_k3'] = {'cost': sd.objective, 'qp': nd, 'ms': 1e3 * td}
            r['same_cost'] = bool(abs(sd.objective - r['nodes_per_rank_1']['cost']) <= 1e-9 * abs(sd.objective))
            r['same_modes'] = bool(np.array_equal(np.array(sd.variables['ub']), np.array(sol.variables['ub'])))
        out['split_frontier_' + name] = r
        del c
    if rank == 0:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'multi_gpu_check.json'), 'w'), indent=1)
        print(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
