#!/usr/bin/env python
"""bench.py -- B&B QP relaxations/s (and ms per MPC step, warm vs cold) of the hybrid-MPC hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1] per instance, configs[2] = 4096 instances over 8 GPUs as the batch):
closed loop of the notebook two-wall cart-pole (T = 20), `--instances` independent initial states per GPU
(warm-start-hybrid-mpc_b200/data/cp20_instances.npy), model error e_t = sigma * randn * x_max, warm-started branch and bound
with tree shifting.  One bench "step" = ONE fused launch (wshmpc_closed_loop) that advances every instance of
the batch by `--window` receding-horizon steps = K3 (device B&B, K1 inside) + K2/K4 (tree shift + plant
update) per MPC step, with no barrier between instances.  Instances are independent: ranks own contiguous
blocks of instances, there is no collective on the data path (weak scaling).

Prints ONE JSON line (rank 0).  `value` = QP relaxations solved by all ranks / max-over-ranks device time,
states resident in HBM; `e2e` = the same loop driven from HOST buffers through the public Python API
(pinned H2D of the measured state and model error, D2H of input, cost and next state, every step).
`--impl reference` times the CPU path (oracle/bnb_ref.py + oracle/qp_core.c, the restatement of the
reference's Python B&B pinned bit-exactly against it -- Gurobi is not available offline) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'bnb_qp_relaxations_per_s'
UNIT = 'QP/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--window', type=int, default=20, help='receding-horizon steps per bench step (one fused launch)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--instances', type=int, default=512, help='independent MPC instances per GPU')
    ap.add_argument('--sigma', type=float, default=0.003, help='model error std (fraction of x_max)')
    ap.add_argument('--max-solves', type=int, default=1024)
    ap.add_argument('--max-roots', type=int, default=512)
    ap.add_argument('--no-extras', action='store_true', help='skip cold / single-instance / cpu legs')
    ap.add_argument('--cpu-seconds', type=float, default=15., help='budget of the cpu_baseline sample')
    return ap.parse_args()


def workload_config(args, world):
    return {'workload': 'cp20_closed_loop_warm_start (two-wall cart-pole, T=20, nx=4, nu=7, 4 binaries/step; '
                        'BASELINE configs[1] per instance, batched as configs[2])',
            'instances_per_gpu': args.instances, 'instances_total': args.instances * world,
            'horizon': 20, 'sigma': args.sigma, 'tol': 0.,
            'window': args.window,
            'step': 'one fused launch = %d receding-horizon steps of every instance: device B&B (K3+K1) + tree shift/plant '
                    'update (K2+K4) per MPC step, task queue over (instance, step), no barrier between instances' % args.window,
            'search': 'best_first / branch_in_time, reference order, no speculative solves',
            'parallelism': 'instances sharded over %d GPU(s), no data-path collective' % world}


def load_instances(lo, hi):
    x = np.load(os.path.join(ROOT, 'warm-start-hybrid-mpc_b200', 'data', 'cp20_instances.npy'))
    idx = np.arange(lo, hi) % len(x)
    return np.ascontiguousarray(x[idx])


def noise(model, n_steps, n_inst, sigma, seed):
    rng = np.random.default_rng(seed)
    return sigma * rng.standard_normal((n_steps, n_inst, model['A'].shape[0])) * model['x_max']


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2.)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [v.strip() for v in s.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU path (oracle restatement of the reference; the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """Closed loop of one instance on one host core.  Returns per-step (t_start, t_end, solves)."""
    k, x0, e, n_steps, budget = job
    from oracle.models import load_model
    from oracle.qp_c import CoreC
    from oracle.bnb_ref import OracleController
    model = load_model('cp20')
    ctl = OracleController(model, CoreC(model, variant=1), hot_start='record')
    x = x0.copy(); ws = None; rows = []
    t_begin = time.perf_counter()
    for t in range(n_steps):
        t0 = time.perf_counter()
        try:
            inc, leaves, solves = ctl.feedforward(x, warm_start=ws)
        except RuntimeError:                     # the CPU QP core gave up on a node: this trajectory ends here
            break
        if inc is None:
            rows.append((t0, time.perf_counter(), solves)); break
        u0 = inc.primal['u'][0]
        ws = ctl.construct_warm_start(leaves, x, u0[:ctl.nuc], u0[ctl.nuc:], e[t])
        x = inc.primal['x'][1] + e[t]
        rows.append((t0, time.perf_counter(), solves))
        if budget is not None and time.perf_counter() - t_begin > budget:
            break
    return rows, ctl.qp_time


def cpu_baseline_sample(args, model, x0, e, seconds):
    """rank 0, N = 1: instance 0 of the workload on ONE host core (the reference loop is single-threaded)."""
    from oracle import qp_c
    qp_c.build()
    t0 = time.perf_counter()
    rows, qp_time = _cpu_worker((0, x0[0], e[:, 0], e.shape[0], seconds))
    wall = time.perf_counter() - t0
    solves = sum(r[2] for r in rows)
    return {'value': solves / wall, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': 'instance 0 of the workload, %d closed-loop steps (1 cold + %d warm), %d QPs in %.1f s; '
                      'oracle/bnb_ref.py + oracle/qp_core.c variant 1 (dual active-set, thin QR, each node started from the dual solution it carries; no pinned-prefix elimination), Python overhead included'
                      % (len(rows), len(rows) - 1, solves, wall),
            'qp_only_value': solves / max(qp_time, 1e-9), 'ms_per_qp': 1e3 * wall / max(solves, 1),
            'host_cores_available': os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = args.gpus
    import multiprocessing as mp
    from oracle import qp_c
    from oracle.models import load_model
    qp_c.build()
    model = load_model('cp20')
    cores = max(1, os.cpu_count() or 1)
    S = args.window
    n_steps = (args.warmup + args.steps) * S
    x0 = load_instances(0, cores)
    e = noise(model, n_steps, cores, args.sigma, 1000)
    jobs = [(k, x0[k], e[:, k], n_steps, None) for k in range(cores)]
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    solves, t_lo, t_hi = 0, [], []
    for rows, _ in res:
        timed = rows[args.warmup * S:]
        if not timed:
            continue
        solves += sum(r[2] for r in timed)
        t_lo.append(timed[0][0]); t_hi.append(timed[-1][1])
    # each worker has its own clock origin (perf_counter is monotonic system-wide on Linux): whole-job time
    elapsed = max(hi - lo for lo, hi in zip(t_lo, t_hi)) if t_lo else float('nan')
    value = solves / elapsed
    cfg = workload_config(args, world)
    sample = ('%d instances (one per host core, fork pool) x %d timed closed-loop MPC steps (= %d bench steps of %d) after %d '
              'warm-up MPC steps (step 0 = cold solve); %d QPs' % (cores, args.steps * S, args.steps, S, args.warmup * S, solves))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * elapsed / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                             'note': 'Gurobi (the reference QP back end) is not installable offline; the reference B&B / '
                                     'warm-start logic is restated in oracle/bnb_ref.py (pinned bit-exactly against the '
                                     'reference code) on the oracle C QP core'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 path
# ---------------------------------------------------------------------------------------------------
def fp64_peak_probe(torch, dev):
    """cuBLAS DGEMM 4096^3, best of 5 (MEASURED_PEAKS.json has no fp64 entry)."""
    n = 4096
    a = torch.randn((n, n), dtype=torch.float64, device=dev); b = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = float('inf')
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a, b); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    del a, b
    return 2. * n ** 3 / (best * 1e-3) / 1e12


def timed_loop(torch, dist, loop, steps, world, body):
    """Times `steps` bench steps bracketed by barrier + synchronize; body(t) launches step t.
    Returns (elapsed ms max over ranks, QPs of all ranks, iterations of all ranks, QPs of this rank); the counter deltas
    of this rank (wshmpc.h d_totals) are left in timed_loop.last."""
    from warm_start_hmpc_b200.closed_loop import reduce_stats
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    before = loop.totals.clone()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(steps):
        body(t)
    e.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = s.elapsed_time(e)
    d = (loop.totals - before).cpu().numpy()
    timed_loop.last = d
    qps, ms_max = reduce_stats(int(d[0]), ms)
    iters, _ = reduce_stats(int(d[1]), ms)
    return ms_max, qps, iters, int(d[0])


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl b200 needs a GPU (there is no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if not os.path.exists(g.LIB):
        if rank == 0:
            g.build()
        if world > 1:
            dist.barrier()
    from warm_start_hmpc_b200.instances import load_model, controller_from_model
    from warm_start_hmpc_b200.closed_loop import ClosedLoop, shard
    model = load_model('cp20')
    ctl = controller_from_model(model, device=local)
    dev = torch.device('cuda', local)
    n_inst, S = args.instances, args.window
    lo, hi = shard(n_inst * world, rank, world)
    x0 = load_instances(lo, hi)
    n_win = args.warmup + args.steps
    e_host = noise(model, (2 * n_win + 4) * S, n_inst, args.sigma, 1000 + rank).reshape(2 * n_win + 4, S, n_inst, -1)
    e_dev = torch.as_tensor(e_host, device=dev)

    loop = ClosedLoop(ctl, n_inst, warm=True, max_solves=args.max_solves, max_roots=args.max_roots)
    tree_bytes = loop.nbytes()
    nx, nu = ctl.mld.nx, ctl.mld.nu
    f64 = dict(dtype=torch.float64, device=dev)
    logs = dict(cost=torch.empty((S, n_inst), **f64), u0=torch.empty((S, n_inst, nu), **f64),
                n_solves=torch.empty((S, n_inst), dtype=torch.int32, device=dev),
                status=torch.empty((S, n_inst), dtype=torch.int32, device=dev))
    loop.reset(x0)
    for w in range(args.warmup):
        loop.run(S, e=e_dev[w], logs=logs)
    torch.cuda.synchronize()

    # ---- device-resident timed region: args.steps fused launches
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.3)
    ev = []

    def dev_step(t):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); loop.run(S, e=e_dev[args.warmup + t], logs=logs); b.record()
        ev.append((a, b))
    l0 = loop.launches
    ms, qps, iters, my_qps = timed_loop(torch, dist, loop, args.steps, world, dev_step)
    cnt = timed_loop.last.astype(float)
    launches = loop.launches - l0
    clocks = sampler.stop()
    ker_ms = [a.elapsed_time(b) for a, b in ev]
    n_active = int(loop.active.sum())
    leaves_mean = float(loop.trees[loop.cur].n_nodes.double().mean())      # warm-start roots per instance after the last shift
    status = loop.out['status'].cpu().numpy()
    value = qps / (ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers: per bench step H2D of the measured states and of the
    # window's model errors (pinned), one fused launch, D2H of the applied inputs, costs and final states
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory()
    xh, eh, uh, ch, xnh = pin(n_inst, nx), pin(S, n_inst, nx), pin(S, n_inst, nu), pin(S, n_inst), pin(n_inst, nx)
    ed = torch.empty((S, n_inst, nx), **f64)
    xh.copy_(loop.x); torch.cuda.synchronize()
    e_host_t = torch.as_tensor(e_host)

    def e2e_step(t):
        eh.copy_(e_host_t[n_win + t])
        loop.x.copy_(xh, non_blocking=True); ed.copy_(eh, non_blocking=True)      # measured states, model errors
        loop.run(S, e=ed, logs=logs)
        uh.copy_(logs['u0'], non_blocking=True); ch.copy_(logs['cost'], non_blocking=True); xnh.copy_(loop.x, non_blocking=True)
        torch.cuda.synchronize()                                                    # the host needs the inputs now
        xh.copy_(xnh)
    ms_e, qps_e, _, _ = timed_loop(torch, dist, loop, args.steps, world, e2e_step)
    e2e = {'value': qps_e / (ms_e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': (S + 1) * n_inst * nx * 8,
           'd2h_bytes_per_step': n_inst * (S * (nu + 1) + nx) * 8, 'ms_per_step': ms_e / args.steps,
           'ms_per_mpc_step_of_the_batch': ms_e / args.steps / S,
           'api': 'ClosedLoop.run(n_steps, e) of warm_start_hmpc_b200 (ctypes -> C ABI wshmpc_closed_loop)'}

    # ---- roofline of the dominant kernel (closed_loop_kernel = K3 with K1 inside + K2/K4)
    pd = ctl.problem
    F_dense = 2. * pd.n ** 2 + 4. * pd.mc * pd.n                     # dense shared-operator form
    # executed form, with the MEASURED mean sizes of this rank's timed solves: a node with d pinned binaries is solved in
    # n_eff = n - d coordinates with a working set of k rows (wshmpc.h d_totals[2], [4], [5])
    my_q = max(cnt[0], 1.)
    kbar, n_eff, k0bar = cnt[2] / my_q, pd.n - cnt[4] / my_q, cnt[5] / my_q
    F_iter = 2. * pd.ns * n_eff + 2. * pd.mc * (pd.nx + pd.nu) + 2. * pd.nb + 4. * n_eff * kbar + kbar ** 2   # pricing + one append
    F_rebuild = k0bar * (4. * n_eff * k0bar / 2. + (k0bar / 2.) ** 2)                                       # k0 appends at growing k
    it_per_qp = iters / max(qps, 1)
    F_qp = it_per_qp * F_iter + F_rebuild
    qp_per_launch = qps / world / max(len(ker_ms), 1)
    it_per_launch = iters / world / max(len(ker_ms), 1)
    ker_avg_ms = float(np.mean(ker_ms))
    achieved = F_qp * qp_per_launch / (ker_avg_ms * 1e-3) / 1e12
    peak = fp64_peak_probe(torch, dev)
    roofline = {'bound': 'tensor', 'kernel': 'closed_loop_kernel', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': achieved / peak, 'traffic': None,
                'peak_source': 'measured in this run: cuBLAS DGEMM 4096^3 best of 5 (fp64 pipe; MEASURED_PEAKS.json has no fp64 entry)',
                'algorithmic': 'flops per QP in the form the kernel executes: iterations x [2 ns n_eff (pricing operator) + 2 mc (nx+nu) '
                               '(stage rows) + 4 n_eff k + k^2 (Gram-Schmidt append, R^-1 column)] + re-factorisation of the k0 inherited '
                               'rows, with the measured means n_eff = n - d = %.1f, k = %.1f, k0 = %.1f, %.1f iterations/QP: %.0f flop/iteration, '
                               '%.0f flop/QP (the dense shared-operator form of SURVEY 8d would be %.0f flop/iteration); %.0f QPs/launch; '
                               'the kernel is latency bound (dependent phases of one small QP per CTA), see DESIGN.md'
                               % (n_eff, kbar, k0bar, it_per_qp, F_iter, F_qp, F_dense, qp_per_launch),
                'achieved_dense_form_tflops': F_dense * it_per_launch / (ker_avg_ms * 1e-3) / 1e12,
                'kernel_ms_avg': ker_avg_ms, 'kernel_share_of_step': float(np.sum(ker_ms) / ms)}
    peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_file):
        roofline['hbm_peak_gbs_measured'] = json.load(open(peaks_file)).get('hbm_gbs')
    # algorithmic HBM bytes of one launch: K1 reads / writes B_node per QP, K2+K4 read and write one dual record
    # per retained leaf and MPC step (the reference shifts every leaf's dual solution, controller.py:431-501)
    B_node = 8 * (2 * pd.nb + pd.nx + pd.n + 2 * pd.m) + 12
    rec_bytes = 8 * loop.h.layout.rec_stride
    roofline['algorithmic_hbm_bytes_per_launch'] = qp_per_launch * B_node + S * n_inst * leaves_mean * 2 * rec_bytes
    roofline['algorithmic_hbm_note'] = ('%.0f QPs x %d B (K1) + %d steps x %d instances x %.1f leaves x 2 x %d B (K2+K4 tree shift)'
                                        % (qp_per_launch, B_node, S, n_inst, leaves_mean, rec_bytes))
    traffic_file = os.path.join(ROOT, 'profiles', 'closed_loop_kernel_traffic.json')
    if os.path.exists(traffic_file):
        roofline['traffic'] = json.load(open(traffic_file)).get('dram_bytes_per_launch')

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, world),
            'roofline': roofline, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
            'ms_per_mpc_step_of_the_batch': ms / args.steps / S,
            'qp_per_mpc_step_per_instance': qps / args.steps / S / (n_inst * world),
            'active_instances_rank0': n_active, 'bnb_status_counts_rank0': {int(k): int((status == k).sum()) for k in np.unique(status)}}
    line['config']['l2'] = 'inputs larger than L2: two device trees of %.1f GB per rank, leaf records touched per step >> 126 MB' % (tree_bytes / 2 / 1e9)

    # ---- extras (outside the timed region): cold start, lock-step API, single instance latency, CPU baseline
    if not args.no_extras:
        del loop
        torch.cuda.empty_cache()
        cold = ClosedLoop(ctl, n_inst, warm=False, max_solves=args.max_solves, max_roots=args.max_roots)
        cold.reset(x0)
        cold.run(1, e=e_dev[0, :1])
        ms_c, qps_c, _, _ = timed_loop(torch, dist, cold, 1, world, lambda t: cold.run(2, e=e_dev[1, :2]))
        line['cold_start'] = {'value': qps_c / (ms_c * 1e-3), 'unit': UNIT, 'ms_per_mpc_step_of_the_batch': ms_c / 2,
                              'qp_per_mpc_step_per_instance': qps_c / 2 / (n_inst * world)}
        line['warm_start'] = {'value': value, 'unit': UNIT, 'ms_per_mpc_step_of_the_batch': ms / args.steps / S,
                              'qp_per_mpc_step_per_instance': line['qp_per_mpc_step_per_instance']}
        del cold
        torch.cuda.empty_cache()
        lock = ClosedLoop(ctl, n_inst, warm=True, max_solves=args.max_solves, max_roots=args.max_roots)
        lock.reset(x0)
        for t in range(3):
            lock.step(e=e_dev[0, t])
        ms_l, qps_l, _, _ = timed_loop(torch, dist, lock, 5, world, lambda t: lock.step(e=e_dev[1, t]))
        line['lock_step_api'] = {'value': qps_l / (ms_l * 1e-3), 'unit': UNIT, 'ms_per_mpc_step_of_the_batch': ms_l / 5,
                                 'note': 'ClosedLoop.step(): one launch pair per MPC step, every instance waits for the slowest one'}
        del lock
        torch.cuda.empty_cache()
        if rank == 0:
            single = {}
            for warm in (True, False):
                L = ClosedLoop(ctl, 1, warm=warm, max_solves=4096, max_roots=512, n_slots=1)
                L.reset(model['x0_nominal'][None])
                L.step(); torch.cuda.synchronize()
                before = L.totals.clone()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k = 5
                s.record()
                for _ in range(k):
                    L.step()
                e.record(); torch.cuda.synchronize()
                d = (L.totals - before).cpu().numpy()
                single['warm' if warm else 'cold'] = {'ms_per_mpc_step': s.elapsed_time(e) / k, 'qp_per_step': float(d[0]) / k,
                                                      'qp_per_s': float(d[0]) / (s.elapsed_time(e) * 1e-3)}
            single['reference_published'] = {'cold_ms_per_mpc_step': 530., 'warm_ms_per_mpc_step': 37.4, 'qp_per_s': 300.,
                                             'source': 'BASELINE.md (Gurobi, unknown CPU)'}
            line['single_instance_nominal'] = single
    if rank == 0 and not args.no_extras:
        # the other BASELINE configs, as information (they are parity-test cases, not the headline): cold batched B&B
        # of the horizon-40 cart-pole (configs[3]) and the root relaxation + dive nodes of the synthetic system (configs[4])
        other = {}
        try:
            m40 = load_model('cp40')
            c40 = controller_from_model(m40, device=local)
            rng = np.random.default_rng(40)
            xs = m40['x0_nominal'][None] + rng.uniform(-1, 1, (148, 4)) * np.array([0.02, 0.01, 0.05, 0.05])
            c40.handle(min(len(xs), c40.default_slots()))
            c40.feedforward_batch(xs[:8], max_solves=2048)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res, _ = c40.feedforward_batch(xs, max_solves=2048)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            ns = res['n_solves'].cpu().numpy(); st = res['status'].cpu().numpy()
            other['cp40_cold_bnb'] = {'instances': int(len(xs)), 'qp': int(ns.sum()), 'qp_per_s': float(ns.sum() / dt),
                                      'wall_ms': 1e3 * dt, 'ms_per_qp_of_the_longest_instance': float(1e3 * dt / max(ns.max(), 1)), 'status_counts': {int(k): int((st == k).sum()) for k in np.unique(st)},
                                      'note': 'horizon 40: n = 280, factor columns beyond the shared-memory budget spill to L2'}
            del c40
            g30 = np.load(os.path.join(ROOT, 'tests', 'golden', 'syn30_nodes.npz'))
            m30 = load_model('syn30')
            c30 = controller_from_model(m30, device=local)
            reps = 5
            N30 = len(g30['status']) * reps
            x30 = np.repeat(g30['x0'][None], N30, 0); lb30 = np.tile(g30['lb'], (reps, 1)); ub30 = np.tile(g30['ub'], (reps, 1))
            h30 = c30.handle(n_slots=min(N30, c30.default_slots()))
            h30.solve_nodes(x30[:8], lb30[:8], ub30[:8]); torch.cuda.synchronize(); t0 = time.perf_counter()
            out = h30.solve_nodes(x30, lb30, ub30)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            other['syn30_k1_dive_nodes'] = {'nodes': int(N30), 'qp_per_s': float(N30 / dt), 'iterations_mean': float(out['iters'].float().mean()),
                                            'note': 'nx = 20, 8 binaries/step, N = 30: n = 360, m = 2640; cold solves of the golden dive nodes'}
            del c30
        except Exception as ex:                     # information only
            other['error'] = repr(ex)
        line['other_configs'] = other
    if rank == 0 and world == 1 and not args.no_extras:
        line['cpu_baseline'] = cpu_baseline_sample(args, model, x0, e_host.reshape(-1, n_inst, nx), args.cpu_seconds)
    elif rank == 0:
        line['cpu_baseline'] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
