"""Multi-GPU plumbing: disorder realisations / initial states / phase-diagram points are independent
(the reference loops over them serially, main.py:467-469), so they are split into contiguous blocks,
one block per rank (one process per GPU), evolved with no communication, and the observable records
are gathered once at the end (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_items, world_size, rank):
    """Contiguous block [lo, hi) of rank ``rank``; sizes differ by at most one, earlier ranks larger."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad rank / world size')
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world_size):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0]
            for r in range(world_size)]


def gather_records(local, n_items, axis=1, group=None, device=None):
    """All-gather per-rank record arrays along the chain axis.

    ``local``: numpy array whose ``axis`` indexes this rank's chains (e.g. Z[T][R_local][L]).
    Returns the full array with ``n_items`` chains on every rank.  Ragged shards are padded to the
    largest shard for the collective and trimmed afterwards.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_items, world)
    big = max(sizes)
    a = np.moveaxis(np.asarray(local), axis, 0)
    if a.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError('local shard has the wrong number of chains')
    pad = np.zeros((big,) + a.shape[1:], dtype=a.dtype)
    pad[:a.shape[0]] = a
    is_complex = np.iscomplexobj(pad)
    t = torch.from_numpy(np.ascontiguousarray(pad.view(np.float64) if is_complex else pad))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    parts = []
    for r, o in enumerate(outs):
        arr = o.cpu().numpy()
        if is_complex:
            arr = arr.view(np.complex128)
        parts.append(arr[:sizes[r]])
    return np.moveaxis(np.concatenate(parts, axis=0), 0, axis)


def disorder_average(local_sum, local_count, group=None, device=None):
    """Sum-reduce partial sums over ranks and divide by the total count (disorder average)."""
    import torch
    import torch.distributed as dist
    s = np.asarray(local_sum, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return s / local_count
    t = torch.from_numpy(np.concatenate([s.reshape(-1), [float(local_count)]]))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu().numpy()
    return t[:-1].reshape(s.shape) / t[-1]


def run_sharded_ensemble(L, J, tau, h_fields_all, n_periods, rank=0, world_size=1, device=0, group=None, **ensemble_kw):
    """The loop the reference runs serially over realisations (main.py:467-469), sharded: rank ``rank`` evolves its
    contiguous block of the chains ``h_fields_all`` [n_items][L] on GPU ``device`` (one FloquetEnsemble, no
    communication), then the records are all-gathered.  Returns Z[T][n_items][L], S_ent, LE, chi on every rank: the
    same arrays, bit for bit, as one GPU evolving all the chains (chains never interact; the per-chain arithmetic does
    not depend on which context or group a chain sits in)."""
    from .engine import FloquetEnsemble
    h = np.atleast_2d(np.asarray(h_fields_all, dtype=float))
    n_items = h.shape[0]
    lo, hi = shard_bounds(n_items, world_size, rank)
    out = {}
    if hi > lo:
        ens = FloquetEnsemble(L, J, tau, h[lo:hi], device=device, **ensemble_kw)
        rec = ens.run(n_periods)
        ens.close()
        out = {k: rec[k] for k in ('Z', 'S_ent', 'LE', 'chi')}
    else:   # more ranks than chains: an empty shard still takes part in the collective
        T = n_periods + 1
        out = {'Z': np.zeros((T, 0, L)), 'S_ent': np.zeros((T, 0, max(L - 1, 0))), 'LE': np.zeros((T, 0)),
               'chi': np.zeros((T, 0, L + 1), dtype=np.int32)}
    if world_size == 1:        # one shard holds everything: no collective (also when a process group exists)
        return out
    dev = f'cuda:{device}' if device is not None else None
    full = {}
    for k, v in out.items():
        a = gather_records(v.astype(np.float64) if k == 'chi' else v, n_items, axis=1, group=group, device=dev)
        full[k] = a.astype(np.int32) if k == 'chi' else a
    return full
