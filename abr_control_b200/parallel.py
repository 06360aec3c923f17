"""Multi-GPU data parallelism for the batched controllers: one process per GPU (torch.distributed, NCCL over NVLink),
the batch split into contiguous row blocks, no collective on the data path.

The states (or trajectories) are independent, so the only exchange is the OPTIONAL all-gather of the
control outputs ``u`` so that every rank holds the full ``(B, n)`` array (BASELINE config 5).  Arm-model and
controller handles are replicated per rank (a few KB of constants).

Two ways to gather: ``all_gather_rows`` (NCCL ``all_gather_into_tensor`` after the kernel) and ``PeerGather`` — the
OSC kernel's own epilogue stores every finished tile into the gathered array of every rank through NVLink peer memory
(CUDA IPC mapped buffers, include/abrb.h ``abrb_gather_*``), so the exchange hides under the arithmetic.
"""
import ctypes as C

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


class _DevView:
    """a raw device pointer as something torch.as_tensor() understands (CUDA array interface)"""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


class PeerGather:
    """Symmetric gather buffers for the fused all-gather epilogue of ``OSC.generate`` (include/abrb.h, abrb_gather_*).

    Every rank allocates ``n_buffers`` gathered ``(rows_total, n_cols)`` arrays, the ranks exchange the CUDA IPC handles
    through ``torch.distributed.all_gather_object`` and map each other's regions.  ``generate(ctrlr, q, dq, target)``
    evaluates this rank's rows and returns the FULL ``(rows_total, n_cols)`` tensor (a view of this rank's current
    buffer) on which the current stream already waits for every peer's rows.  Buffers alternate call by call.
    """

    def __init__(self, rows_total, n_cols, dtype, group=None, n_buffers=2):
        from . import _lib

        self._L = _lib.lib()
        self._check = _lib.check
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.rows_total, self.n_cols, self.dtype = int(rows_total), int(n_cols), dtype
        self.n_buffers = int(n_buffers)
        itemsize = torch.empty((), dtype=dtype).element_size()
        h = C.c_void_p()
        self._check(self._L.abrb_gather_create(self.rank, self.world, self.rows_total * self.n_cols * itemsize,
                                               self.n_buffers, C.byref(h)))
        self._h = h
        mine = C.create_string_buffer(64)
        self._check(self._L.abrb_gather_export(h, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=group)
        for r, raw in enumerate(handles):
            if r != self.rank:
                self._check(self._L.abrb_gather_import(h, r, C.create_string_buffer(raw, 64)))
        typestr = {torch.float64: "<f8", torch.float32: "<f4"}[dtype]
        self._views = []
        for i in range(self.n_buffers):
            ptr = self._L.abrb_gather_buffer(h, i)
            self._views.append(torch.as_tensor(_DevView(ptr, (self.rows_total, self.n_cols), typestr, self),
                                               device=torch.device("cuda", torch.cuda.current_device())))
        self._next = 0
        dist.barrier(group=group)  # nobody stores into a peer before that peer has finished mapping

    def generate(self, ctrlr, q, dq, target, row0=None, **kw):
        """this rank's rows ``q, dq, target`` (CUDA tensors) -> the gathered (rows_total, n) result of all ranks"""
        i = self._next
        self._next = (i + 1) % self.n_buffers
        if row0 is None:
            row0 = shard_range(self.rows_total, self.rank, self.world)[0]
        ctrlr._generate_gather(q, dq, target, self._h, i, int(row0), **kw)
        self._check(self._L.abrb_gather_wait(self._h, torch.cuda.current_stream(q.device).cuda_stream))
        return self._views[i]

    def status(self):
        return self._L.abrb_gather_status(self._h)

    def close(self):
        if self._h is not None:
            self._views = []
            dist.barrier(group=self.group)  # every rank has stopped storing into its peers
            self._L.abrb_gather_destroy(self._h)
            self._h = None
