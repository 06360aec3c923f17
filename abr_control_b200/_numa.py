"""Host-side placement helper for the *_host entry points.

Page-locked buffers live on the NUMA node of the core that allocated and first touched them; if that is not the node the
GPU hangs off, every H2D/D2H copy crosses the inter-socket link.  ``gpu_local_cpus(device)`` runs the enclosed block on
the cores NVML reports as closest to the GPU, so buffers created inside end up GPU-local; the previous affinity is
restored on exit (worker threads started later are not restricted)."""
import contextlib
import os


@contextlib.contextmanager
def gpu_local_cpus(device_index=0, uuid=None):
    old = None
    try:
        import pynvml

        pynvml.nvmlInit()
        h = None
        if uuid:
            for cand in (f"GPU-{uuid}", str(uuid)):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
                    break
                except Exception:
                    h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        old = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)  # calling thread -> the GPU's ideal CPU set
    except Exception:
        old = None
    try:
        yield
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass
