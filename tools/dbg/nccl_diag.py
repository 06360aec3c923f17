#!/usr/bin/env python
"""Multi-GPU plumbing diagnostic (torchrun, one rank per GPU): topology, peer access, NCCL transport and the
device-event time of the collectives bench.py uses, legacy CUDA IPC (the route of the fused peer-store epilogue)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
res = {"rank": rank, "world": world}
if rank == 0:
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, file=sys.stderr)
    except Exception as e:  # pragma: no cover
        print("topo failed", e, file=sys.stderr)
res["can_access_peer"] = [bool(torch.cuda.can_device_access_peer(local, j)) for j in range(torch.cuda.device_count()) if j != local]
dist.init_process_group("nccl", device_id=dev)


def ev_time(fn, n, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3  # us


for name, shape, dtype in (("ag_u_f64_B65536", (65536, 6), torch.float64), ("ag_u_f32_B131072", (131072, 6), torch.float32),
                           ("ag_64MB", (16 << 20,), torch.float32)):
    x = torch.ones(shape, dtype=dtype, device=dev)
    out = torch.empty((world * shape[0],) + tuple(shape[1:]), dtype=dtype, device=dev)
    us = ev_time(lambda: dist.all_gather_into_tensor(out, x), 200 if x.numel() < (4 << 20) else 20)
    res[name] = {"us": us, "bytes_per_rank": x.numel() * x.element_size(),
                 "bus_GBps": x.numel() * x.element_size() * (world - 1) / us / 1e3}
# wall-clock view (what a host-timed loop sees): includes the launch path
x = torch.ones((65536, 6), dtype=torch.float64, device=dev)
out = torch.empty((world * 65536, 6), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize()
res["ag_u_f64_wall_us"] = (time.perf_counter() - t0) / 100 * 1e6

print("partial", json.dumps(res), file=sys.stderr, flush=True)
# legacy CUDA IPC: every rank cudaMalloc's 8 MB, exports the handle, opens all peers' and times a peer write
rt = C.CDLL("libcudart.so.12")
ptr = C.c_void_p()
assert rt.cudaMalloc(C.byref(ptr), C.c_size_t(8 << 20)) == 0
class IpcHandle(C.Structure):  # cudaIpcMemHandle_t is passed BY VALUE to cudaIpcOpenMemHandle
    _fields_ = [("raw", C.c_char * 64)]


handle = IpcHandle()
rc = rt.cudaIpcGetMemHandle(C.byref(handle), ptr)
res["ipc_get"] = rc
handles = [None] * world
dist.all_gather_object(handles, bytes(bytearray(handle)))
peers = []
for r in range(world):
    if r == rank:
        peers.append(ptr.value)
        continue
    h = IpcHandle.from_buffer_copy(handles[r])
    p = C.c_void_p()
    rc = rt.cudaIpcOpenMemHandle(C.byref(p), h, C.c_uint(1))
    if rc:
        rt.cudaGetLastError()
    res.setdefault("ipc_open", []).append(rc)
    peers.append(p.value)
if all(v == 0 for v in res.get("ipc_open", [])) and world > 1:
    src = torch.ones(3 << 18, dtype=torch.float64, device=dev)  # 6 MB
    nxt = peers[(rank + 1) % world]
    s = torch.cuda.current_stream().cuda_stream

    def push():
        rt.cudaMemcpyAsync(C.c_void_p(nxt), C.c_void_p(src.data_ptr()), C.c_size_t(src.numel() * 8), C.c_int(3), C.c_void_p(s))

    us = ev_time(push, 100)
    res["ipc_peer_copy_6MB"] = {"us": us, "GBps": src.numel() * 8 / us / 1e3}
dist.barrier()
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    print(json.dumps(allres))
dist.destroy_process_group()
