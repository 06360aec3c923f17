"""GPU suite, two or more GPUs of one box: the sharded controller and the fused all-gather epilogue (the OSC kernel
stores its rows into every rank's gathered array over NVLink peer memory, abr_control_b200/parallel.py PeerGather)
against a single-GPU evaluation of the whole batch and against NCCL's all-gather.  Skipped on a one-GPU box; the host
logic of the N > 1 path is covered on CPU by tests/test_distributed.py (gloo)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok, why = True, ""
    try:
        from abr_control_b200 import parallel
        from abr_control_b200.arms import jaco2, ur5
        from abr_control_b200.controllers import OSC, AvoidObstacles, Damping

        rng = np.random.default_rng(17)
        for arm, dtype, B, kw in (
                (ur5, torch.float64, 4099, dict(kp=10, ctrlr_dof=[True] * 6, use_C=True)),
                (jaco2, torch.float32, 10000, dict(kp=200, vmax=[0.5, 0], ctrlr_dof=[True, True, True, False, False, False]))):
            rc = arm.Config()
            nulls = None
            if arm is jaco2:  # BASELINE config 5's controller
                nulls = [AvoidObstacles(rc, obstacles=[[0.09596, -0.2661, 0.64204, 0.05]], threshold=0.2), Damping(rc, kv=10)]
            ctrlr = OSC(rc, null_controllers=nulls, **kw)
            q, dq, tg = (torch.as_tensor(rng.uniform(0, 2 * np.pi, (B, 6)), device=dev, dtype=dtype),
                         torch.as_tensor(rng.uniform(0, 3, (B, 6)), device=dev, dtype=dtype),
                         torch.as_tensor(rng.uniform(-1, 1, (B, 6)), device=dev, dtype=dtype))
            whole = ctrlr.generate(q, dq, tg)  # every rank evaluates the whole batch itself: the reference result
            lo, hi = parallel.shard_range(B, rank, world)
            pg = parallel.PeerGather(B, 6, dtype)
            for rep in range(5):  # buffers alternate; results must not depend on which one is used
                full = pg.generate(ctrlr, q[lo:hi].contiguous(), dq[lo:hi].contiguous(), tg[lo:hi].contiguous())
                torch.cuda.synchronize()
                if not torch.equal(full, whole):
                    ok, why = False, f"fused gather differs ({arm.__name__}, rep {rep}): {(full - whole).abs().max().item()}"
            if pg.status() != 0:
                ok, why = False, "a gather wait timed out"
            nccl = parallel.all_gather_rows(ctrlr.generate(q[lo:hi], dq[lo:hi], tg[lo:hi]), B)
            if not torch.equal(nccl, whole):
                ok, why = False, "NCCL gather differs"
            sc = parallel.ShardedController(ctrlr)
            if not torch.equal(sc.generate(q, dq, tg, gather=True), whole):
                ok, why = False, "ShardedController(gather=True) differs"
            pg.close()
    except Exception as e:  # pragma: no cover
        ok, why = False, repr(e)
    q_out.put((rank, ok, why))
    dist.barrier()
    dist.destroy_process_group()


def test_fused_peer_gather_matches_single_gpu_and_nccl():
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok, _ in res), res
