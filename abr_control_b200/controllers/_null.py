"""Common part of the secondary ("null space") controllers: standalone batched ``generate(q, dq)``."""
import ctypes as C
import weakref

import numpy as np

from .. import _lib
from . import _batch
from .controller import Controller


class NullController(Controller):
    def __init__(self, robot_config):
        super().__init__(robot_config)
        self._owners = weakref.WeakSet()  # OSC instances that embedded our parameters (weak: they may die first)

    def _params(self):
        raise NotImplementedError

    def _dirty(self):
        for o in list(self._owners):
            o._invalidate()

    def __setattr__(self, name, value):
        # the reference reads a secondary controller's attributes on every call; an OSC that embedded them in its
        # native handle has to rebuild it when one changes (damping.kv = ..., avoid.obstacles = ...)
        object.__setattr__(self, name, value)
        if not name.startswith("_") and name != "robot_config" and "_owners" in self.__dict__:
            self._dirty()

    def generate(self, q, dq):
        """(n,) -> (n,) float64;  (B,n) -> (B,n) (NumPy in/out or CUDA tensor in/out)."""
        rc = self.robot_config
        qa, dqa, single, kind, f32 = _batch.prep_state(rc, q, dq)
        p = self._params()
        L = _lib.lib()
        B = qa.shape[0]
        if kind == "torch":
            import torch

            with torch.cuda.device(qa.device):
                u = torch.empty_like(qa)
                fn = L.abrb_null_generate_f32 if f32 else L.abrb_null_generate_f64
                _lib.check(fn(rc.handle, C.byref(p), qa.data_ptr(), dqa.data_ptr(), u.data_ptr(), B,
                              torch.cuda.current_stream(qa.device).cuda_stream))
            return u[0] if single else u
        import torch  # host path of the standalone call goes through pinned staging tensors

        dev = torch.device("cuda", torch.cuda.current_device())
        tq = torch.as_tensor(qa).to(dev)
        tdq = torch.as_tensor(dqa).to(dev)
        tu = torch.empty_like(tq)
        fn = L.abrb_null_generate_f32 if f32 else L.abrb_null_generate_f64
        _lib.check(fn(rc.handle, C.byref(p), tq.data_ptr(), tdq.data_ptr(), tu.data_ptr(), B,
                      torch.cuda.current_stream(dev).cuda_stream))
        u = tu.cpu().numpy()
        return np.array(u[0], dtype=np.float64) if single else u
