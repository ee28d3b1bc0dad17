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
