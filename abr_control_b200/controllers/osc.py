"""Batched operational-space controller.

Same constructor and ``generate`` signature as ``abr_control.controllers.OSC``
(/root/reference/abr_control/controllers/osc.py:53-66 and :217-219).  ``generate`` accepts one state
(``q`` of shape ``(n,)`` -> fresh writable float64 ``(n,)`` array, like the reference) or a batch
(``(B, n)`` NumPy -> NumPy, CUDA tensor -> CUDA tensor).  One call is one fused kernel launch:
chain walk, J / M / g / C dq, Cholesky of M, task-space inertia with the reference's det-threshold / pinv
branch (osc.py:120-147), task PD with optional velocity limiting (osc.py:198-215), orientation error by either
algorithm (osc.py:149-196), gravity / Coriolis compensation and the null-space filter for the secondary
controllers (osc.py:310-318).
"""
import ctypes as C

import numpy as np

from .. import _abi, _lib
from . import _batch
from ._null import NullController
from .controller import Controller


class OSC(Controller):
    def __init__(
        self,
        robot_config,
        kp=1,
        ko=None,
        kv=None,
        ki=0,
        vmax=None,
        ctrlr_dof=None,
        null_controllers=None,
        use_g=True,
        use_C=False,
        orientation_algorithm=0,
    ):
        super().__init__(robot_config)
        self.kp = kp
        self.ko = kp if ko is None else ko
        self.kv = np.sqrt(self.kp + self.ko) if kv is None else kv
        self.ki = ki
        self.null_controllers = null_controllers
        self.use_g = use_g
        self.use_C = use_C
        self.orientation_algorithm = orientation_algorithm
        if self.ki != 0:
            raise NotImplementedError(
                "ki != 0 keeps per-controller integrator state (osc.py:81-82, :262-264); not supported in the batched engine"
            )
        if ctrlr_dof is None:
            ctrlr_dof = [True, True, True, False, False, False]
        self.ctrlr_dof = np.copy(ctrlr_dof)
        self.n_ctrlr_dof = np.sum(self.ctrlr_dof)
        self.task_space_gains = np.array([self.kp] * 3 + [self.ko] * 3)
        self.lamb = self.task_space_gains / self.kv
        if self.n_ctrlr_dof > robot_config.N_JOINTS:
            print(
                f"\nRobot has fewer DOF ({robot_config.N_JOINTS}) than the specified number of "
                f"space dimensions to control ({self.n_ctrlr_dof}), Poor performance may result.\n"
            )
        self.vmax = vmax
        if vmax is not None:
            self.sat_gain_xyz = vmax[0] / self.kp * self.kv
            self.sat_gain_abg = vmax[1] / self.ko * self.kv
            self.scale_xyz = vmax[0] / self.kp * self.kv
            self.scale_abg = vmax[1] / self.ko * self.kv
        if self.orientation_algorithm not in (0, 1):
            raise Exception(
                f"Invalid algorithm number {self.orientation_algorithm} for calculating orientation error"
            )
        for nc in self.null_controllers or []:
            if not isinstance(nc, NullController):
                raise TypeError("null_controllers must be abr_control_b200 Damping / RestingConfig / AvoidObstacles")
            nc._owners.append(self)
        self.training_signal = None
        # the reference stores `training_signal` on every call (osc.py:297); for large host batches it doubles the
        # device-to-host traffic, so throughput-minded callers may switch it off
        self.record_training_signal = True
        self._options = {}
        self._handle = None

    # ------------------------------------------------------------------ native handle
    def _invalidate(self):
        if self._handle is not None:
            _lib.lib().abrb_osc_destroy(self._handle)
            self._handle = None

    def _native(self):
        if self._handle is None:
            n = self.robot_config.N_JOINTS
            nulls = [nc._params() for nc in (self.null_controllers or [])]
            p = _abi.osc_params(
                n, kp=self.kp, ko=self.ko, kv=self.kv, ki=self.ki, vmax=self.vmax, ctrlr_dof=list(self.ctrlr_dof),
                null=nulls, use_g=self.use_g, use_C=self.use_C, orientation_algorithm=self.orientation_algorithm,
            )
            h = C.c_void_p()
            _lib.check(_lib.lib().abrb_osc_create(self.robot_config.handle, C.byref(p), C.byref(h)))
            self._handle = h
            for name, value in self._options.items():
                _lib.check(_lib.lib().abrb_osc_set_option(h, name.encode(), value))
        return self._handle

    def set_option(self, name, value):
        """Execution option of the native controller (include/abrb.h, abrb_osc_set_option), e.g.
        ``set_option("two_launch_min", 131072)``.  Kept across parameter changes."""
        self._options[name] = float(value)
        if self._handle is not None:
            _lib.check(_lib.lib().abrb_osc_set_option(self._handle, name.encode(), float(value)))

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    # ------------------------------------------------------------------ generate
    def generate(self, q, dq, target, target_velocity=None, ref_frame="EE", xyz_offset=None):
        rc = self.robot_config
        qa, dqa, single, kind, f32 = _batch.prep_state(rc, q, dq)
        B = qa.shape[0]
        tgt, tstride = _batch.prep_rows(target, qa, kind, 6, "target")
        tv, tvstride = (None, 0)
        if target_velocity is not None:
            tv, tvstride = _batch.prep_rows(target_velocity, qa, kind, 6, "target_velocity")
        fid = rc.frame_id(ref_frame)
        xo = None
        if xyz_offset is not None and not np.allclose(np.asarray(xyz_offset, dtype=float), 0):
            xo = (C.c_double * 3)(*[float(v) for v in np.asarray(xyz_offset, dtype=float).reshape(3)])
        L = _lib.lib()
        h = self._native()
        if kind == "torch":
            import torch

            with torch.cuda.device(qa.device):
                u = torch.empty_like(qa)
                tr = torch.empty_like(qa)
                fn = L.abrb_osc_generate_f32 if f32 else L.abrb_osc_generate_f64
                _lib.check(fn(h, fid, xo, qa.data_ptr(), dqa.data_ptr(), tgt.data_ptr(), tstride, _batch.ptr(tv),
                              tvstride, u.data_ptr(), tr.data_ptr(), B,
                              torch.cuda.current_stream(qa.device).cuda_stream))
        else:
            u = _batch.host_out(qa.shape, qa.dtype)
            tr = _batch.host_out(qa.shape, qa.dtype) if (self.record_training_signal or single) else None
            fn = L.abrb_osc_generate_host_f32 if f32 else L.abrb_osc_generate_host_f64
            _lib.check(fn(h, fid, xo, qa.ctypes.data, dqa.ctypes.data, tgt.ctypes.data, tstride, _batch.ptr(tv),
                          tvstride, u.ctypes.data, _batch.ptr(tr), B))
        if single:
            self.training_signal = np.array(tr[0], dtype=np.float64) if kind == "numpy" else tr[0]
            return np.array(u[0], dtype=np.float64) if kind == "numpy" else u[0]
        self.training_signal = tr
        return u

    def generate_into(self, q, dq, target, u_out, training_out=None, target_velocity=None, ref_frame="EE",
                      xyz_offset=None):
        """Allocation-free batched ``generate`` for hot loops: all arguments are contiguous CUDA tensors of one dtype
        (``q, dq, u_out, training_out``: (B, n); ``target, target_velocity``: (B, 6) or (6,)); the result is written
        into ``u_out`` on the current torch stream.  No shape massaging, one ctypes call."""
        import torch

        rc = self.robot_config
        B, n = q.shape
        if n != rc.N_JOINTS or dq.shape != q.shape or u_out.shape != q.shape or not q.is_cuda:
            raise ValueError("generate_into: q, dq, u_out must be CUDA tensors of shape (B, n_joints)")
        f32 = q.dtype == torch.float32
        for t in (dq, target, u_out, training_out, target_velocity):
            if t is not None and (t.dtype != q.dtype or not t.is_contiguous() or t.device != q.device):
                raise ValueError("generate_into: tensors must share dtype/device and be contiguous")
        L = _lib.lib()
        fn = L.abrb_osc_generate_f32 if f32 else L.abrb_osc_generate_f64
        xo = None
        if xyz_offset is not None:
            xo = (C.c_double * 3)(*[float(v) for v in xyz_offset])
        _lib.check(fn(self._native(), rc.frame_id(ref_frame), xo, q.data_ptr(), dq.data_ptr(), target.data_ptr(),
                      6 if target.dim() == 2 else 0, None if target_velocity is None else target_velocity.data_ptr(),
                      0 if target_velocity is None or target_velocity.dim() == 1 else 6, u_out.data_ptr(),
                      None if training_out is None else training_out.data_ptr(), B,
                      torch.cuda.current_stream(q.device).cuda_stream))
        return u_out

    def rollout(self, q, dq, target, steps, dt=1e-3, ref_frame="EE", xyz_offset=None, record=("q", "dq", "u")):
        """Closed-loop rollout on the GPU: ``steps`` iterations of
        ``u = generate(q, dq, target); ddq = M^-1 (u + g - C dq); dq += ddq dt; q += dq dt``
        (semi-implicit Euler, as /root/reference/abr_control/arms/twojoint/arm_sim.py:131-132).

        Returns ``(q_final, dq_final, traj)`` with ``traj[k]`` of shape ``(steps, B, n)`` for k in ``record``.
        CUDA tensors in -> CUDA tensors out; NumPy in -> NumPy out (staged through the current CUDA device).
        """
        import torch

        rc = self.robot_config
        qa, dqa, single, kind, f32 = _batch.prep_state(rc, q, dq)
        if kind == "numpy":
            dev = torch.device("cuda", torch.cuda.current_device())
            qa, dqa = torch.as_tensor(qa).to(dev), torch.as_tensor(dqa).to(dev)
        else:
            qa, dqa = qa.clone(), dqa.clone()
        tgt, tstride = _batch.prep_rows(target, qa, "torch", 6, "target")
        B, n = qa.shape
        fid = rc.frame_id(ref_frame)
        xo = None
        if xyz_offset is not None and not np.allclose(np.asarray(xyz_offset, dtype=float), 0):
            xo = (C.c_double * 3)(*[float(v) for v in np.asarray(xyz_offset, dtype=float).reshape(3)])
        traj = {k: torch.empty((steps, B, n), dtype=qa.dtype, device=qa.device) for k in record}
        L = _lib.lib()
        fn = L.abrb_osc_rollout_f32 if f32 else L.abrb_osc_rollout_f64
        with torch.cuda.device(qa.device):
            _lib.check(fn(self._native(), fid, xo, qa.data_ptr(), dqa.data_ptr(), tgt.data_ptr(), tstride, int(steps),
                          float(dt), _batch.ptr(traj.get("q")), _batch.ptr(traj.get("dq")), _batch.ptr(traj.get("u")),
                          B, torch.cuda.current_stream(qa.device).cuda_stream))
        if kind == "numpy":
            qa, dqa = qa.cpu().numpy(), dqa.cpu().numpy()
            traj = {k: v.cpu().numpy() for k, v in traj.items()}
        if single:
            return qa[0], dqa[0], {k: v[:, 0] for k, v in traj.items()}
        return qa, dqa, traj
