"""Sliding-mode controller (Slotine & Li).

Reference: /root/reference/abr_control/controllers/sliding.py:6-99 —
``u = M ddq_ref + C dq_ref + g - kd s`` with ``s = dq - dq_ref``; in Cartesian mode the references come from
``pinv(J[:3])`` of the position error, in joint mode from the joint error.  ``self.s`` (the signal the reference hands
to its dynamics adaptation) is stored on every call.  One state or a batch, one kernel launch per call.
"""
import ctypes as C

import numpy as np

from .. import _lib
from . import _batch
from .controller import Controller


class Sliding(Controller):
    def __init__(self, robot_config, kd=160.0, lamb=30.0, cartesian=True):
        super().__init__(robot_config)
        self.kd = kd
        self.lamb = lamb
        self.cartesian = cartesian
        self.s = None

    def generate(self, q, dq, target, target_velocity=0, target_acc=0, ref_frame="EE", offset=None):
        import torch

        rc = self.robot_config
        n = rc.N_JOINTS
        w = 3 if self.cartesian else n
        qa, dqa, single, kind, f32 = _batch.prep_state(rc, q, dq)
        B = qa.shape[0]

        def rows(x, what):
            if x is None or (np.isscalar(x) and x == 0):
                return None, 0
            if np.isscalar(x):
                x = np.full(w, float(x))
            return _batch.prep_rows(x, qa, kind, w, what)

        tgt, ts = _batch.prep_rows(target, qa, kind, w, "target")
        tv, tvs = rows(target_velocity, "target_velocity")
        ta, tas = rows(target_acc, "target_acc")
        xo = None
        if offset is not None and not np.allclose(np.asarray(offset, dtype=float), 0):
            xo = (C.c_double * 3)(*[float(v) for v in np.asarray(offset, dtype=float).reshape(3)])
        fid = rc.frame_id(ref_frame)
        arrays = [qa, dqa, tgt, tv, ta]
        if kind == "numpy":
            dev = torch.device("cuda", torch.cuda.current_device())
            arrays = [None if a is None else torch.as_tensor(a).to(dev) for a in arrays]
        tq = arrays[0]
        u, s = torch.empty_like(tq), torch.empty_like(tq)
        L = _lib.lib()
        fn = L.abrb_sliding_generate_f32 if f32 else L.abrb_sliding_generate_f64
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        with torch.cuda.device(tq.device):
            _lib.check(fn(rc.handle, float(self.kd), float(self.lamb), int(bool(self.cartesian)), fid, xo, ptr(arrays[0]),
                          ptr(arrays[1]), ptr(arrays[2]), ts, ptr(arrays[3]), tvs, ptr(arrays[4]), tas, u.data_ptr(),
                          s.data_ptr(), B, torch.cuda.current_stream(tq.device).cuda_stream))
        if kind == "numpy":
            u, s = u.cpu().numpy(), s.cpu().numpy()
            self.s = np.array(s[0], dtype=np.float64) if single else s
            return np.array(u[0], dtype=np.float64) if single else u
        self.s = s[0] if single else s
        return u[0] if single else u
