"""Multi-GPU plumbing: the CV units (iteration, fold) are independent, so they are
partitioned across ranks with no data-path collective; the only communication is one
all_gather of fixed-size per-unit records at the end.  Works with
the ``nccl`` backend on GPUs and with ``gloo`` on CPU tensors (tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_iterations(n_iter, rank, world):
    """Contiguous block of CV iterations for ``rank`` (iterations of one block share the
    fold-invariant work, so whole iterations stay on one rank)."""
    base, rem = divmod(n_iter, world)
    lo = rank * base + min(rank, rem)
    return list(range(lo, lo + base + (1 if rank < rem else 0)))


def gather_records(records, device=None):
    """records: (n_local, width) int32 array, rows = ``[unit_id, ...payload]``.  Returns the
    rows of all ranks sorted by unit id (every rank gets the same array)."""
    rec = np.ascontiguousarray(records, dtype=np.int32).reshape(len(records), -1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec[np.argsort(rec[:, 0], kind='stable')]
    world = dist.get_world_size()
    dev = device if device is not None else torch.device('cpu')
    n = torch.tensor([rec.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = int(max(c.item() for c in counts))
    pad = np.full((nmax, rec.shape[1]), -1, dtype=np.int32)
    pad[:rec.shape[0]] = rec
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    rows = np.concatenate([o.cpu().numpy()[:int(c.item())] for o, c in zip(outs, counts)])
    return rows[np.argsort(rows[:, 0], kind='stable')]


def init_from_env(device_index=None):
    """Joins the process group a launcher (torchrun) described in the environment, once.  Returns
    ``(rank, world, local_rank)``; a plain single-process run returns ``(0, 1, 0)`` without
    touching torch.distributed.  Backend: ``CPSD_DIST_BACKEND`` if set, else nccl when every rank
    owns its own GPU, else gloo (several ranks sharing one GPU, or CPU-only tests)."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size(), int(os.environ.get('LOCAL_RANK', '0'))
    if world <= 1:
        return 0, 1, 0
    local = int(os.environ.get('LOCAL_RANK', '0'))
    backend = os.environ.get('CPSD_DIST_BACKEND')
    ngpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if backend is None:
        backend = 'nccl' if ngpu >= int(os.environ.get('LOCAL_WORLD_SIZE', world)) else 'gloo'
    if ngpu:
        torch.cuda.set_device(local % ngpu if device_index is None else device_index)
    if backend == 'nccl':
        dist.init_process_group('nccl', device_id=torch.device('cuda', local % ngpu))
    else:
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size(), local


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def comm_device():
    """Device the gather tensors live on: the current GPU under nccl, the host under gloo."""
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def shard_units(n_units, group, rank=None, world=None):
    """Indices of the units owned by ``rank`` when consecutive blocks of ``group`` units (the
    folds of one CV iteration / one subsample) are dealt to the ranks in contiguous runs."""
    if rank is None or world is None:
        rank, world = rank_world()
    n_groups = -(-n_units // group)
    mine = shard_iterations(n_groups, rank, world)
    return [u for g in mine for u in range(g * group, min((g + 1) * group, n_units))]


def gather_predictions(unit_ids, preds, device=None, enabled=True):
    """The one collective of a sharded run: every rank contributes ``preds[i]`` (1-D int label
    array, any length) for unit ``unit_ids[i]``; every rank gets ``{unit_id: labels}`` for all
    units.  Records are ``[unit_id, n, labels..., padding]`` int32 rows."""
    rank, world = rank_world()
    if world == 1 or not enabled:
        return {int(u): np.asarray(p) for u, p in zip(unit_ids, preds)}
    dev = device if device is not None else comm_device()
    wmax = torch.tensor([max([len(p) for p in preds] + [0])], dtype=torch.int64, device=dev)
    dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    w = int(wmax.item())
    rec = np.full((len(unit_ids), 2 + w), -1, dtype=np.int32)
    for i, (u, p) in enumerate(zip(unit_ids, preds)):
        p = np.asarray(p, dtype=np.int32)
        rec[i, 0], rec[i, 1] = u, len(p)
        rec[i, 2:2 + len(p)] = p
    rows = gather_records(rec, device=dev)
    return {int(r[0]): r[2:2 + int(r[1])].copy() for r in rows}


def gather_objects(unit_ids, objs):
    """Small picklable per-unit objects (e.g. the winning parameter dict of a nested search) from
    all ranks, as a list ordered by unit id (every rank gets it)."""
    rank, world = rank_world()
    pairs = list(zip([int(u) for u in unit_ids], objs))
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, pairs)
        pairs = [p for part in allp for p in part]
    return [o for _, o in sorted(pairs, key=lambda t: t[0])]
