"""Multi-GPU data parallelism for the batched controllers: one process per GPU (torch.distributed, NCCL over NVLink),
the batch split into contiguous row blocks, no collective on the data path.

The states (or trajectories) are independent, so the only exchange is the OPTIONAL all-gather of the
control outputs ``u`` so that every rank holds the full ``(B, n)`` array (BASELINE config 5).  Arm-model and
controller handles are replicated per rank (a few KB of constants).
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_range(B, rank, world):
    """Rows [lo, hi) of a B-row batch owned by ``rank``: contiguous blocks, sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world or B < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(B, world):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def all_gather_rows(local, B, group=None):
    """All-gather row blocks produced under ``shard_range`` into the full ``(B, ...)`` tensor on every rank.

    Equal shards use one ``all_gather_into_tensor`` (NCCL ring/NVLS over NVSwitch); ragged shards are padded to the
    largest shard and trimmed.
    """
    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    rank = dist.get_rank(group)
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = torch.empty((B,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    m = max(sizes)
    pad = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * m,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)], dim=0)


class ShardedController:
    """Evaluate a batched controller on this rank's row block of a global batch.

    ``generate(q, dq, target, ..., gather=False)`` takes the GLOBAL arrays (host NumPy or tensors), evaluates rows
    ``shard_range(B, rank, world)`` on this rank's GPU and returns the local block, or — with ``gather=True`` — the
    full ``(B, n)`` result on every rank.
    """

    def __init__(self, controller, group=None):
        self.controller = controller
        self.group = group

    def _world(self):
        if dist is not None and dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def generate(self, q, dq, target, gather=False, **kw):
        rank, world = self._world()
        B = len(q)
        lo, hi = shard_range(B, rank, world)
        tgt = target[lo:hi] if np.ndim(target) == 2 else target
        for name in ("target_velocity", "target_acc"):  # per-state rows follow the shard, broadcast rows do not
            v = kw.get(name)
            if v is not None and np.ndim(v) == 2:
                kw = dict(kw, **{name: v[lo:hi]})
        u = self.controller.generate(q[lo:hi], dq[lo:hi], tgt, **kw)
        if not gather or world == 1:
            return u
        was_numpy = isinstance(u, np.ndarray)
        t = torch.as_tensor(u)
        if dist.get_backend(self.group) == "nccl" and not t.is_cuda:
            t = t.cuda()
        full = all_gather_rows(t, B, self.group)
        return full.cpu().numpy() if was_numpy else full
