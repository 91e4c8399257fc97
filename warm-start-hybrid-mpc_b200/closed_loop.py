"""Batched closed loop of independent MPC instances, entirely on the device.

Restates the experiment loop of notebooks/cart_pole_with_walls/statistical_analysis.py:93-196
(without its Gurobi legs) for `n_inst` instances at once: per step  branch and bound (K3, warm- or
cold-started)  ->  construct_warm_start + plant update x <- x_1|t + e_t (K2 + K4).  Two device trees
are used alternately (the shift reads one and writes the other); nothing synchronises with the host
unless the caller reads a result.  Instances are independent: under torch.distributed each rank owns
a contiguous block of instances (`shard`) and no collective is on the data path.
"""
import numpy as np


def shard(n_total, rank, world):
    """Contiguous block [lo, hi) of the instances owned by `rank`."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ClosedLoop(object):

    def __init__(self, controller, n_inst, warm=True, tol=0., max_solves=1024, max_roots=512, n_slots=None):
        import torch
        self.ctl, self.n_inst, self.warm, self.tol = controller, n_inst, warm, tol
        self.max_solves, self.max_roots = max_solves, max_roots
        slots = n_slots if n_slots is not None else min(n_inst, controller.default_slots())
        self.h = h = controller.handle(slots)
        dev = h.torch_device
        cap_nodes, cap_recs = max_roots + 2 * max_solves + 2, max_roots + max_solves + 1
        self.trees = [h.new_tree(n_inst, cap_nodes, cap_recs) for _ in range(2)]
        self.cur = 0
        nx, nu = controller.mld.nx, controller.mld.nu
        f64 = dict(dtype=torch.float64, device=dev)
        self.xbuf = torch.zeros((2, n_inst, nx), **f64)     # states, ping-pong (the shift writes the other half)
        self.xpar = 0
        self.u0 = torch.zeros((n_inst, nu), **f64)
        self.active = torch.ones(n_inst, dtype=torch.int32, device=dev)
        self.out = dict(cost=torch.zeros(n_inst, **f64), node=torch.zeros(n_inst, dtype=torch.int32, device=dev),
                        primal=torch.zeros((n_inst, h.layout.primal), **f64),
                        n_solves=torch.zeros(n_inst, dtype=torch.int32, device=dev),
                        status=torch.zeros(n_inst, dtype=torch.int32, device=dev))
        self.totals = torch.zeros(8, dtype=torch.int64, device=dev)     # wshmpc.h d_totals: QP solves, iterations, working-set statistics
        self.gid = torch.arange(n_inst, dtype=torch.int64, device=dev)     # global identity of the instance in every slot (rebalance moves instances)
        self.fresh = True
        self.launches = 0
        self.events = None          # list -> (start, after K3, after K2+K4) CUDA events of every step

    @property
    def x(self):
        """Current states [n_inst, nx] (CUDA tensor view)."""
        return self.xbuf[self.xpar]

    def nbytes(self):
        return sum(t.nbytes() for t in self.trees)

    def reset(self, x0):
        """New initial states [n_inst, nx]; the next step starts from the root node."""
        import torch
        self.x.copy_(torch.as_tensor(np.asarray(x0, dtype=float) if not torch.is_tensor(x0) else x0))
        self.active.fill_(1)
        self.totals.zero_()
        self.fresh = True

    def rebalance(self, group=None, min_gap=2):
        """Periodic load balancing between the ranks of a torch.distributed job (rebalance.py): live instances move from
        the busiest ranks into the slots of dead instances of the idlest ones, with their warm-start trees.  Call it
        between steps / windows.  Returns (sent, received, plan)."""
        from .rebalance import rebalance
        return rebalance(self.trees[self.cur], self.x, self.active, self.gid, group, min_gap)

    def step(self, e=None, x=None):
        """One receding-horizon step of every instance.  `x` (optional, [n_inst, nx] CUDA tensor)
        overrides the state (measured state fed back from the host); `e` is the model error e_t."""
        h = self.h
        if x is not None:
            self.x.copy_(x)
        tree = self.trees[self.cur]
        if self.fresh or not self.warm:
            h.tree_init_root(tree); self.launches += 1
            self.fresh = False
        if self.events is not None:
            import torch
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        h.bnb_solve(self.x, tree, tol=self.tol, max_solves=self.max_solves, active=self.active, out=self.out,
                    totals=self.totals)
        self.launches += 1
        if self.events is not None:
            ev[1].record()
        # K2 + K4 (in cold mode only its plant update matters: the next step re-initialises the root)
        new = self.trees[1 - self.cur]
        h.shift_tree(self.x, e, tree, self.out['cost'], self.out['primal'], new, active=self.active,
                     x_next=self.xbuf[1 - self.xpar], u0=self.u0)
        self.cur = 1 - self.cur
        self.launches += 1
        if self.events is not None:
            ev[2].record()
            self.events.append(tuple(ev))
        self.xpar = 1 - self.xpar
        return self.out

    def step_with_plant(self, plant):
        """One receding-horizon step against a TRUE plant instead of the linear model + noise: branch and bound (K3), the
        applied input goes to `plant(x [n, nx], u0 [n, nu]) -> next measured state [n, nx]` (host numpy, e.g.
        plants.CartPoleWithWalls.simulate), the model error e_t = x_measured - x_1|t is what the warm start of the next
        step is corrected with (construct_warm_start's e0, controller.py:503-564), then K2 + K4.
        Returns (out, u0 [n, nu] numpy, e [n, nx] numpy)."""
        import torch
        h = self.h
        tree = self.trees[self.cur]
        if self.fresh or not self.warm:
            h.tree_init_root(tree); self.launches += 1
            self.fresh = False
        h.bnb_solve(self.x, tree, tol=self.tol, max_solves=self.max_solves, active=self.active, out=self.out, totals=self.totals)
        self.launches += 1
        nx, nu, T = self.ctl.mld.nx, self.ctl.mld.nu, self.ctl.T
        prim = self.out['primal'].cpu().numpy()
        x_now = self.x.cpu().numpy()
        u0 = prim[:, (T + 1) * nx:(T + 1) * nx + nu]
        x_pred = prim[:, nx:2 * nx]
        live = (self.active.cpu().numpy() != 0) & np.isfinite(self.out['cost'].cpu().numpy())
        x_meas = np.array(x_now)
        if live.any():
            x_meas[live] = plant(x_now[live], u0[live])
        e = np.where(live[:, None], x_meas - x_pred, 0.)
        ed = torch.as_tensor(e, device=self.x.device)
        new = self.trees[1 - self.cur]
        h.shift_tree(self.x, ed, tree, self.out['cost'], self.out['primal'], new, active=self.active,
                     x_next=self.xbuf[1 - self.xpar], u0=self.u0)
        if live.any():
            # the next step starts from the MEASURED state itself (x_1|t + (x_measured - x_1|t) is one rounding away from it)
            lv = torch.as_tensor(live, device=self.x.device)
            self.xbuf[1 - self.xpar][lv] = torch.as_tensor(x_meas[live], device=self.x.device)
        self.cur = 1 - self.cur
        self.xpar = 1 - self.xpar
        self.launches += 1
        return self.out, u0, e

    def run(self, n_steps, e=None, logs=None):
        """`n_steps` receding-horizon steps of every instance in ONE launch (wshmpc_closed_loop): the fused
        form of calling step() n_steps times, without a barrier between the steps of different instances.
        e: [n_steps, n_inst, nx] CUDA tensor of model errors or None.  Returns per-step logs
        (cost, u0, n_solves, status), each [n_steps, n_inst, ...]; self.x holds the final states."""
        h = self.h
        if self.warm:
            par, fresh = self.cur, self.fresh
        else:
            par, fresh = self.cur, True
        # states and trees flip together: make the state parity equal to the tree parity
        if self.xpar != par:
            self.xbuf[par].copy_(self.xbuf[self.xpar]); self.xpar = par
        logs = h.closed_loop(n_steps, self.warm, fresh, par, self.xbuf, e, self.active, self.trees, self.out,
                             tol=self.tol, max_solves=self.max_solves, totals=self.totals, logs=logs)
        self.launches += 2
        self.fresh = False
        self.cur = (self.cur + n_steps) & 1
        self.xpar = (self.xpar + n_steps) & 1
        return logs


    def run_mailbox(self, n_steps, plant, timeout_s=120., logs=None):
        """`n_steps` receding-horizon steps of every instance with the HOST in the loop every step and no barrier between
        instances (wshmpc_closed_loop with a wshmpc_mailbox): one persistent launch solves; this thread polls the pinned
        mailbox, and whenever instances have published a step it calls
            plant(idx [k] instance indices, step [k], u0 [k, nu], x_pred [k, nx]) -> x_measured [k, nx]  or  (x_measured, e)
        (host numpy; u0 is NaN where a step has no incumbent) and hands the measured states and the model errors
        e = x_measured - x_pred back -- the per-step use of `feedforward` + `construct_warm_start`
        (statistical_analysis.py:120-196) for a batch of independent plants.  Returns the device logs of run() plus
        'host' = dict(u0, x1, cost, status) numpy arrays [n_steps, n_inst, ...] as the host saw them."""
        import time
        import torch
        from .capi import Mailbox
        h, N = self.h, self.n_inst
        if getattr(self, '_mailbox', None) is None:
            self._mailbox = Mailbox(h, N)
        mb = self._mailbox
        mb.clear()
        if self.warm:
            par, fresh = self.cur, self.fresh
        else:
            par, fresh = self.cur, True
        if self.xpar != par:
            self.xbuf[par].copy_(self.xbuf[self.xpar]); self.xpar = par
        nx, nu = self.ctl.mld.nx, self.ctl.mld.nu
        host = dict(u0=np.full((n_steps, N, nu), np.nan), x1=np.zeros((n_steps, N, nx)), cost=np.full((n_steps, N), np.inf),
                    status=np.zeros((n_steps, N), dtype=np.int32))
        logs = h.closed_loop(n_steps, self.warm, fresh, par, self.xbuf, None, self.active, self.trees, self.out,
                             tol=self.tol, max_solves=self.max_solves, totals=self.totals, logs=logs, mailbox=mb)
        self.launches += 2
        seen = np.zeros(N, dtype=np.int32)
        t_end = time.time() + timeout_s
        n_left = N * n_steps
        try:
            while n_left > 0:
                idx = np.nonzero(mb.out_step > seen)[0]
                if idx.size == 0:
                    if time.time() > t_end:
                        raise RuntimeError('run_mailbox: the kernel published nothing for %.0f s' % timeout_s)
                    continue
                s = seen[idx]
                u0 = mb.out_u0[idx].copy(); x1 = mb.out_x1[idx].copy()
                host['u0'][s, idx] = u0; host['x1'][s, idx] = x1
                host['cost'][s, idx] = mb.out_cost[idx]; host['status'][s, idx] = mb.out_status[idx]
                ans = plant(idx, s, u0, x1)
                if isinstance(ans, tuple):
                    xm, e = ans
                else:
                    xm = np.asarray(ans, dtype=float); e = xm - x1
                live = np.isfinite(u0).all(axis=1)
                mb.in_x[idx] = np.where(live[:, None], xm, x1)
                mb.in_e[idx] = np.where(live[:, None], e, 0.)
                mb.in_step[idx] = s + 1            # after the data (x86 stores are not reordered)
                seen[idx] = s + 1
                n_left -= idx.size
                t_end = time.time() + timeout_s
        except BaseException:
            mb.stop[0] = 1                          # never leave the kernel waiting for an answer
            torch.cuda.synchronize(h.torch_device)
            raise
        self.fresh = False
        self.cur = (self.cur + n_steps) & 1
        self.xpar = (self.xpar + n_steps) & 1
        logs['host'] = host
        return logs


def reduce_stats(n_units, elapsed_ms, device=None):
    """Whole-job aggregate under torch.distributed (one process per GPU, no data-path collective):
    units are summed over ranks, time is the MAX over ranks.  Works with NCCL (CUDA tensors) and gloo
    (CPU tensors); without an initialised process group it returns the inputs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(n_units), float(elapsed_ms)
    dev = device if device is not None else ('cuda' if dist.get_backend() == 'nccl' else 'cpu')
    u = torch.tensor([int(n_units)], dtype=torch.int64, device=dev)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(u.item()), float(t.item())
