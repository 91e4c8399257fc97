"""Periodic load balancing of closed-loop instances between the GPUs of a job (north_star item 4).

Instances are independent, so the data path needs no collective; what drifts is the LOAD: trajectories that become
infeasible leave the loop (controller returns None, statistical_analysis.py:99-108), and a rank whose instances died
idles while the others still work.  `rebalance` evens the number of live instances out:

  1. all ranks exchange their live counts (one small all-gather),
  2. every rank derives the same greedy plan (busiest -> idlest until the spread is <= 1, bounded by free slots),
  3. each move ships the instance's state -- measured state x, its identity (global id), and the warm-start tree of the
     next step (node arrays + the dual records of its leaves, ~1 MB on the T = 20 cart-pole) -- as two packed tensors
     over torch.distributed point-to-point (NCCL: NVLink peer copies; gloo in the CPU tests),
  4. the receiver unpacks it into the slot of a dead instance; the sender marks its slot dead.

The payload is the tree the NEXT search starts from (after the shift every leaf owns record number = its node number),
so a moved instance continues exactly as it would have at home: results are identical to the unbalanced run
(tests/test_distributed_cpu.py checks the round trip bit for bit).
"""
import torch


def plan_moves(live, free):
    """Greedy plan from the live counts and free slots of all ranks: list of (src, dst, count).  Deterministic, so every
    rank computes the same plan from the same all-gathered counts."""
    live, free = list(live), list(free)
    moves = []
    while True:
        src = max(range(len(live)), key=lambda r: (live[r], -r))
        dst = min(range(len(live)), key=lambda r: (live[r], r))
        gap = live[src] - live[dst]
        n = min(gap // 2, free[dst])
        if n <= 0:
            break
        moves.append((src, dst, n))
        live[src] -= n; live[dst] += n; free[dst] -= n; free[src] += n
    # merge repeated pairs
    merged = {}
    for s, d, n in moves:
        merged[(s, d)] = merged.get((s, d), 0) + n
    return [(s, d, n) for (s, d), n in merged.items()]


def pack(tree, x, gid, idx):
    """State of the instances `idx` (LongTensor) of a device tree -> (header [3] int64, ints [m, ...] int32,
    doubles [m, ...] float64).  Only the live part of every array travels (nodes < n_nodes, records < n_recs)."""
    m = idx.numel()
    nn = int(tree.n_nodes[idx].max()) if m else 0
    nr = int(tree.n_recs[idx].max()) if m else 0
    W = tree.words
    ints = torch.zeros((m, 2 + nn * (3 + 2 * W)), dtype=torch.int32, device=x.device)
    ints[:, 0] = tree.n_nodes[idx]; ints[:, 1] = tree.n_recs[idx]
    o = 2
    for name in ('depth', 'alive', 'rec'):
        ints[:, o:o + nn] = getattr(tree, name)[idx, :nn]; o += nn
    ints[:, o:o + nn * W] = tree.bits[idx, :nn].reshape(m, nn * W); o += nn * W
    ints[:, o:o + nn * W] = tree.mask[idx, :nn].reshape(m, nn * W)
    nx = x.shape[1]
    S = tree.rec_dual.shape[2]
    dbl = torch.zeros((m, 1 + nx + nn + nr * (1 + S)), dtype=torch.float64, device=x.device)
    dbl[:, 0] = gid[idx].to(torch.float64)
    dbl[:, 1:1 + nx] = x[idx]
    o = 1 + nx
    dbl[:, o:o + nn] = tree.lb[idx, :nn]; o += nn
    dbl[:, o:o + nr] = tree.rec_dobj[idx, :nr]; o += nr
    dbl[:, o:] = tree.rec_dual[idx, :nr].reshape(m, nr * S)
    hdr = torch.tensor([m, nn, nr], dtype=torch.int64, device=x.device)
    return hdr, ints, dbl


def unpack(tree, x, gid, slots, hdr, ints, dbl):
    """Inverse of pack: writes the travelling instances into the local slots `slots` (LongTensor)."""
    m, nn, nr = [int(v) for v in hdr]
    if nn > tree.cap_nodes or nr > tree.cap_recs:
        raise RuntimeError('rebalance: an incoming tree (%d nodes, %d records) does not fit the local capacity (%d, %d)'
                           % (nn, nr, tree.cap_nodes, tree.cap_recs))
    W = tree.words
    tree.n_nodes[slots] = ints[:, 0]; tree.n_recs[slots] = ints[:, 1]
    o = 2
    for name in ('depth', 'alive', 'rec'):
        getattr(tree, name)[slots, :nn] = ints[:, o:o + nn]; o += nn
    tree.bits[slots, :nn] = ints[:, o:o + nn * W].reshape(m, nn, W); o += nn * W
    tree.mask[slots, :nn] = ints[:, o:o + nn * W].reshape(m, nn, W)
    nx = x.shape[1]
    S = tree.rec_dual.shape[2]
    gid[slots] = dbl[:, 0].to(gid.dtype)
    x[slots] = dbl[:, 1:1 + nx]
    o = 1 + nx
    tree.lb[slots, :nn] = dbl[:, o:o + nn]; o += nn
    tree.rec_dobj[slots, :nr] = dbl[:, o:o + nr]; o += nr
    tree.rec_dual[slots, :nr] = dbl[:, o:].reshape(m, nr, S)


def rebalance(tree, x, active, gid, group=None, min_gap=2):
    """Evens out the live instances of all ranks of `group` (see module docstring).  tree: the tree the next search
    starts from; x [n_inst, nx], active [n_inst] int32, gid [n_inst] int64: all modified in place.
    Returns (instances sent, instances received, plan)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0, 0, []
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = x.device
    mine = torch.tensor([int(active.sum()), int(active.numel() - active.sum())], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    live = [int(c[0]) for c in allc]; free = [int(c[1]) for c in allc]
    if max(live) - min(live) < min_gap:
        return 0, 0, []
    plan = plan_moves(live, free)
    sent = got = 0
    for src, dst, n in plan:
        if rank == src:
            idx = torch.nonzero(active)[:, 0][-n:]                     # the last n live instances leave
            hdr, ints, dbl = pack(tree, x, gid, idx)
            dist.send(hdr, dst, group=group)
            dist.send(ints, dst, group=group); dist.send(dbl, dst, group=group)
            active[idx] = 0
            tree.n_nodes[idx] = 0; tree.n_recs[idx] = 0
            sent += n
        elif rank == dst:
            hdr = torch.zeros(3, dtype=torch.int64, device=dev)
            dist.recv(hdr, src, group=group)
            m, nn, nr = [int(v) for v in hdr]
            ints = torch.zeros((m, 2 + nn * (3 + 2 * tree.words)), dtype=torch.int32, device=dev)
            dbl = torch.zeros((m, 1 + x.shape[1] + nn + nr * (1 + tree.rec_dual.shape[2])), dtype=torch.float64, device=dev)
            dist.recv(ints, src, group=group); dist.recv(dbl, src, group=group)
            slots = torch.nonzero(active == 0)[:, 0][:m]
            unpack(tree, x, gid, slots, hdr, ints, dbl)
            active[slots] = 1
            got += m
    return sent, got, plan
