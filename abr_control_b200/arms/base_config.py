"""Batched ``robot_config``: the reference's arm-model duck type, evaluated on the GPU.

Mirrors the public methods of ``abr_control.arms.base_config.BaseConfig``
(/root/reference/abr_control/arms/base_config.py:210-415) — same names, same argument order:

    g(q)  dJ(name, q, dq, x=None)  J(name, q, x=None)  M(q)  R(name, q)  quaternion(name, q)
    C(q, dq)  T(name, q)  Tx(name, q, x=None)  T_inv(name, q, x=None)

Each accepts
  * ONE state — ``q`` of shape ``(n,)`` (list / NumPy): returns exactly the reference's shapes and dtypes
    (``J, dJ, M, g, C, R`` rounded to float32 as base_config.py:223,247,270,285,301,336 do; ``Tx, T, T_inv``
    float64), as fresh writable ndarrays;
  * a BATCH — ``q`` of shape ``(B, n)``: NumPy in -> NumPy out (host buffers, copies inside the call), or a CUDA
    ``torch.Tensor`` in -> CUDA tensors out on the current torch stream (no host round trip).  Batched results
    keep the compute dtype (float64 by default, float32 for float32 inputs) and are stacked along axis 0.

All arithmetic runs in hand-written sm_100a kernels behind ``libabrb.so``; there is no SymPy, no code
generation, no cache directory and no CPU fallback.
"""
import ctypes as C

import numpy as np

from .. import _abi, _lib
from ..controllers._batch import host_out as _host_out

try:  # torch is only needed when the caller hands in CUDA tensors
    import torch
except Exception:  # pragma: no cover
    torch = None

_RBD_KEYS = ("Tx", "T", "R", "T_inv", "quat", "J", "dJ", "M", "g", "C")


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


class BaseConfig:
    """Batched arm model built from a flat chain descriptor (see ``abr_control_b200/arms/data/*.json``).

    Parameters
    ----------
    desc : dict
        chain descriptor: n_joints, n_links, L0, A, B, E (3x4 blocks), link_inertia, gravity
    dtype : numpy dtype, optional (Default: float64)
        compute precision used for host (NumPy / list) inputs; CUDA tensors use their own dtype
    """

    def __init__(self, desc, ROBOT_NAME="robot", dtype=np.float64, **kwargs):
        kwargs.pop("use_cython", None)  # accepted for signature compatibility (base_config.py:78); meaningless here
        if kwargs:
            raise TypeError(f"unexpected arguments {sorted(kwargs)}")
        self.desc = desc
        self.ROBOT_NAME = ROBOT_NAME
        self.N_JOINTS = int(desc["n_joints"])
        self.N_LINKS = int(desc["n_links"])
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("dtype must be float32 or float64")
        self._cdesc = _abi.chain_desc_from_dict(desc)
        self._M_LINKS = [np.diag(row) for row in np.asarray(desc["link_inertia"], dtype=float)]
        self._M_JOINTS = [np.zeros((6, 6)) for _ in range(self.N_JOINTS)]
        self.L = np.asarray(desc.get("L", []), dtype=float)
        self.START_ANGLES = np.asarray(desc.get("start_angles", np.zeros(self.N_JOINTS)), dtype=float)
        self.x_zeros = np.zeros(3)
        self._handle = C.c_void_p()
        _lib.check(_lib.lib().abrb_model_create(C.byref(self._cdesc), C.byref(self._handle)))
        self._frame_ids = {}

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().abrb_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def handle(self):
        return self._handle

    @property
    def is_orthonormal(self):
        return bool(_lib.lib().abrb_model_is_orthonormal(self._handle))

    def frame_id(self, name):
        """Frame name -> id.  Unknown names raise like the reference (arms/ur5/config.py:336-337)."""
        fid = self._frame_ids.get(name)
        if fid is None:
            fid = _lib.lib().abrb_frame_id(self._handle, str(name).encode())
            if fid < 0:
                raise Exception(f"Invalid transformation name: {name}")
            self._frame_ids[name] = fid
        return fid

    def _shapes(self):
        n = self.N_JOINTS
        return dict(Tx=(3,), T=(4, 4), R=(3, 3), T_inv=(4, 4), quat=(4,), J=(6, n), dJ=(6, n), M=(n, n), g=(n,),
                    C=(n, n))

    def _prep(self, arr, dtype=None):
        """-> (array as (B,n) contiguous, single?, kind) where kind is 'torch' or 'numpy'."""
        n = self.N_JOINTS
        if _is_torch(arr):
            if not arr.is_cuda:
                raise ValueError("torch inputs must be CUDA tensors (use NumPy for host data)")
            if arr.dtype not in (torch.float32, torch.float64):
                raise ValueError("torch inputs must be float32 or float64")
            single = arr.dim() == 1
            a = arr.reshape(1, -1) if single else arr
            if a.dim() != 2 or a.shape[1] != n:
                raise ValueError(f"expected shape ({n},) or (B, {n}), got {tuple(arr.shape)}")
            return a.contiguous(), single, "torch"
        a = np.asarray(arr, dtype=self.dtype if dtype is None else dtype)
        single = a.ndim == 1
        a = a.reshape(1, -1) if single else a
        if a.ndim != 2 or a.shape[1] != n:
            raise ValueError(f"expected shape ({n},) or (B, {n}), got {np.shape(arr)}")
        return np.ascontiguousarray(a), single, "numpy"

    def eval(self, q, dq=None, name="EE", x=None, want=("J", "M", "g")):
        """Evaluate several quantities in ONE kernel launch.  Returns a dict keyed like ``want``.

        This is the batched superset of the reference's one-quantity-per-call methods; the individual
        methods below call it with a single key.
        """
        want = tuple(want)
        for k in want:
            if k not in _RBD_KEYS:
                raise KeyError(k)
        qa, single, kind = self._prep(q, np.float64 if np.ndim(q) == 1 and not _is_torch(q) else None)
        need_dq = ("dJ" in want) or ("C" in want)
        dqa = None
        if dq is not None:
            dqa, _, kind2 = self._prep(dq, qa.dtype if kind == "numpy" else None)
            if kind2 != kind or dqa.shape != qa.shape or (kind == "torch" and dqa.dtype != qa.dtype):
                raise ValueError("q and dq must have the same type, dtype and shape")
        elif need_dq:
            raise ValueError("dq is required for dJ / C")
        fid = self.frame_id(name)
        xo = None
        if x is not None and not np.allclose(np.asarray(x, dtype=float), 0):
            xo = (C.c_double * 3)(*[float(v) for v in np.asarray(x, dtype=float).reshape(3)])
        B = qa.shape[0]
        shapes = self._shapes()
        out = _abi.RbdOut()
        res = {}
        L = _lib.lib()
        if kind == "torch":
            f32 = qa.dtype == torch.float32
            with torch.cuda.device(qa.device):
                for k in want:
                    res[k] = torch.empty((B,) + shapes[k], dtype=qa.dtype, device=qa.device)
                    setattr(out, k, res[k].data_ptr())
                fn = L.abrb_rbd_eval_f32 if f32 else L.abrb_rbd_eval_f64
                stream = torch.cuda.current_stream(qa.device).cuda_stream
                _lib.check(fn(self._handle, fid, xo, qa.data_ptr(), dqa.data_ptr() if dqa is not None else None, B,
                              C.byref(out), stream))
        else:
            f32 = qa.dtype == np.float32
            for k in want:
                res[k] = _host_out((B,) + shapes[k], qa.dtype)
                setattr(out, k, res[k].ctypes.data)
            fn = L.abrb_rbd_eval_host_f32 if f32 else L.abrb_rbd_eval_host_f64
            _lib.check(fn(self._handle, fid, xo, qa.ctypes.data, dqa.ctypes.data if dqa is not None else None, B,
                          C.byref(out)))
        if single:
            res = {k: v[0] for k, v in res.items()}
        return res

    def eval_into(self, q, dq, out, name="EE", x=None):
        """Allocation-free batched evaluation for hot loops: ``q``/``dq`` contiguous CUDA tensors (B, n) of one dtype,
        ``out`` a dict key -> preallocated contiguous CUDA tensor of the right shape (keys as in ``eval``)."""
        B = q.shape[0]
        if not q.is_cuda or q.dim() != 2 or q.shape[1] != self.N_JOINTS or not q.is_contiguous():
            raise ValueError("eval_into: q must be a contiguous CUDA tensor of shape (B, n_joints)")
        if dq is not None and (dq.shape != q.shape or dq.dtype != q.dtype or dq.device != q.device or not dq.is_contiguous()):
            raise ValueError("eval_into: dq must match q")
        o = _abi.RbdOut()
        shapes = self._shapes()
        for k, t in out.items():
            if k not in _RBD_KEYS or t.dtype != q.dtype or not t.is_contiguous() or t.device != q.device:
                raise ValueError(f"eval_into: bad output {k}")
            if tuple(t.shape) != (B,) + shapes[k]:
                raise ValueError(f"eval_into: output {k} must have shape {(B,) + shapes[k]}")
            setattr(o, k, t.data_ptr())
        xo = None
        if x is not None:
            xo = (C.c_double * 3)(*[float(v) for v in x])
        L = _lib.lib()
        fn = L.abrb_rbd_eval_f32 if q.dtype == torch.float32 else L.abrb_rbd_eval_f64
        with torch.cuda.device(q.device):  # the launch goes to the tensors' device, whatever the current one is
            _lib.check(fn(self._handle, self.frame_id(name), xo, q.data_ptr(), None if dq is None else dq.data_ptr(), B,
                          C.byref(o), torch.cuda.current_stream(q.device).cuda_stream))
        return out

    def _one(self, key, q, dq=None, name="EE", x=None, ref32=False):
        v = self.eval(q, dq=dq, name=name, x=x, want=(key,))[key]
        if isinstance(v, np.ndarray) and v.ndim == len(self._shapes()[key]):
            # single state: the reference's dtype contract
            return np.array(v, dtype="float32") if ref32 else np.array(v, dtype=np.float64)
        return v

    # ------------------------------------------------------------------ the reference's public surface
    def g(self, q):
        """Joint-space gravity force (base_config.py:210-223)."""
        return self._one("g", q, ref32=True)

    def dJ(self, name, q, dq, x=None):
        """Time derivative of the Jacobian (base_config.py:225-247)."""
        return self._one("dJ", q, dq=dq, name=name, x=x, ref32=True)

    def J(self, name, q, x=None):
        """6 x n Jacobian of point ``x`` in frame ``name`` (base_config.py:249-270)."""
        return self._one("J", q, name=name, x=x, ref32=True)

    def M(self, q):
        """Joint-space inertia matrix (base_config.py:272-285)."""
        return self._one("M", q, ref32=True)

    def R(self, name, q):
        """Rotation matrix of frame ``name`` (base_config.py:287-301)."""
        return self._one("R", q, name=name, ref32=True)

    def quaternion(self, name, q):
        """Unit quaternion (w, x, y, z) of frame ``name`` (base_config.py:304-318)."""
        return self._one("quat", q, name=name)

    def C(self, q, dq):
        """Centrifugal/Coriolis matrix such that ``C @ dq`` is the force (base_config.py:320-336)."""
        return self._one("C", q, dq=dq, ref32=True)

    def T(self, name, q):
        """4 x 4 transform of frame ``name`` (base_config.py:338-369)."""
        return self._one("T", q, name=name)

    def Tx(self, name, q, x=None):
        """World position of point ``x`` of frame ``name`` (base_config.py:371-392)."""
        return self._one("Tx", q, name=name, x=x)

    def T_inv(self, name, q, x=None):
        """Inverse transform [[R^T, -R^T t], [0, 1]] (base_config.py:394-415; ``x`` is unused there too)."""
        return self._one("T_inv", q, name=name)


def builtin_config(arm, **kwargs):
    """Config for one of the arms shipped with the reference: 'ur5', 'jaco2', 'threejoint', 'twojoint'."""
    return BaseConfig(_abi.load_arm_json(arm), ROBOT_NAME=arm, **kwargs)
