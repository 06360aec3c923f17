"""CPU suite: the N>1 host logic (sharding + optional all-gather of u) with gloo, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from abr_control_b200 import parallel


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 7, 8, 65536, 1048576 + 3):
        for world in (1, 2, 3, 4, 8):
            rs = [parallel.shard_range(B, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == parallel.shard_sizes(B, world)
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


class _FakeController:
    """stands in for OSC on a CPU-only box: u = f(q, dq, target) row-wise, so shards can be checked exactly"""

    def generate(self, q, dq, target, **kw):
        q, dq, target = (np.asarray(a, dtype=np.float64) for a in (q, dq, target))
        tv = kw.get("target_velocity")
        u = 2.0 * q - 0.5 * dq + target[..., :1]
        return u + (0 if tv is None else np.asarray(tv)[..., :1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, ragged_ok, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        qq, dq = rng.normal(size=(B, 6)), rng.normal(size=(B, 6))
        tgt, tv = rng.normal(size=(B, 6)), rng.normal(size=(B, 6))
        ctrl = parallel.ShardedController(_FakeController())
        full_ref = _FakeController().generate(qq, dq, tgt, target_velocity=tv)
        lo, hi = parallel.shard_range(B, rank, world)
        local = ctrl.generate(qq, dq, tgt, target_velocity=tv)
        ok = np.array_equal(local, full_ref[lo:hi])
        gathered = ctrl.generate(qq, dq, tgt, gather=True, target_velocity=tv)
        ok = ok and gathered.shape == full_ref.shape and np.array_equal(gathered, full_ref)
        bc = ctrl.generate(qq, dq, tgt[0], gather=True)  # broadcast target row
        ok = ok and np.array_equal(bc, _FakeController().generate(qq, dq, tgt[0]))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 64), (2, 65), (3, 100)])
def test_sharded_generate_and_all_gather_gloo(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, True, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world)) and all(ok for _, ok in res)
