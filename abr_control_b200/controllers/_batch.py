"""Shared input/output plumbing for the batched controllers (host NumPy buffers or CUDA torch tensors)."""
import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def prep_state(rc, q, dq):
    """-> (q2, dq2, single, kind, f32) with q2/dq2 contiguous (B, n)."""
    qa, single, kind = rc._prep(q, np.float64 if (np.ndim(q) == 1 and not is_torch(q)) else None)
    dqa, _, kind2 = rc._prep(dq, qa.dtype if kind == "numpy" else None)
    if kind != kind2 or qa.shape != dqa.shape or (kind == "torch" and qa.dtype != dqa.dtype):
        raise ValueError("q and dq must have the same type, dtype and shape")
    f32 = (qa.dtype == torch.float32) if kind == "torch" else (qa.dtype == np.float32)
    return qa, dqa, single, kind, f32


def prep_rows(x, like, kind, width, what):
    """(width,) -> broadcast row (stride 0); (B, width) -> per-state rows (stride width)."""
    B = like.shape[0]
    if kind == "torch":
        t = x if is_torch(x) else torch.as_tensor(np.asarray(x, dtype=np.float64), device=like.device)
        t = t.to(device=like.device, dtype=like.dtype)
        if t.dim() == 1 and t.shape[0] == width:
            return t.contiguous(), 0
        if t.dim() == 2 and t.shape == (B, width):
            return t.contiguous(), width
    else:
        a = np.asarray(x.detach().cpu() if is_torch(x) else x, dtype=like.dtype)
        if a.ndim == 1 and a.shape[0] == width:
            return np.ascontiguousarray(a), 0
        if a.ndim == 2 and a.shape == (B, width):
            return np.ascontiguousarray(a), width
    raise ValueError(f"{what} must have shape ({width},) or ({B}, {width})")


def ptr(a):
    if a is None:
        return None
    return a.data_ptr() if is_torch(a) else a.ctypes.data


def host_out(shape, dtype):
    """NumPy result buffer for the host-pointer entry points.  From 64 KiB up it is page-locked (torch's caching host
    allocator: cheap after the first call, returned to the cache when the array is dropped) so that the library's
    device->host copy is a plain DMA instead of a staged copy through the driver's bounce buffer."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes >= (1 << 16):
        import torch

        t = torch.empty(tuple(int(s) for s in shape), dtype=torch.float32 if np.dtype(dtype) == np.float32 else torch.float64,
                        pin_memory=True)
        return t.numpy()  # keeps `t` alive
    return np.empty(shape, dtype=dtype)

