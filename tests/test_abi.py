"""CPU suite: the C-ABI boundary (include/abrb.h <-> libabrb.so <-> ctypes mirror).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from abr_control_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "abrb.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abrb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libabrb.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.abrb_version() == 200


def test_struct_layouts_match_the_header():
    """sizeof/offsetof as gcc sees include/abrb.h == the ctypes mirror."""
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "abrb.h"
int main(void){
  printf("%zu %zu %zu %zu\n", sizeof(abrb_chain_desc), sizeof(abrb_rbd_out), sizeof(abrb_null_params), sizeof(abrb_osc_params));
  printf("%zu %zu %zu %zu\n", offsetof(abrb_chain_desc, E), offsetof(abrb_chain_desc, gravity), offsetof(abrb_null_params, obstacles), offsetof(abrb_osc_params, null));
  printf("%zu %zu %zu\n", offsetof(abrb_osc_params, ctrlr_dof), offsetof(abrb_osc_params, n_null), offsetof(abrb_null_params, threshold));
  return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    got = [int(v) for v in out]
    want = [C.sizeof(_abi.ChainDesc), C.sizeof(_abi.RbdOut), C.sizeof(_abi.NullParams), C.sizeof(_abi.OscParams),
            _abi.ChainDesc.E.offset, _abi.ChainDesc.gravity.offset, _abi.NullParams.obstacles.offset,
            _abi.OscParams.null.offset, _abi.OscParams.ctrlr_dof.offset, _abi.OscParams.n_null.offset,
            _abi.NullParams.threshold.offset]
    assert got == want


def test_model_handles_and_error_codes_without_compute():
    lib = _lib.lib()
    cd = _abi.chain_desc_from_dict(_abi.load_arm_json("ur5"))
    h = C.c_void_p()
    assert lib.abrb_model_create(C.byref(cd), C.byref(h)) == 0
    assert lib.abrb_model_n_joints(h) == 6 and lib.abrb_model_is_orthonormal(h) == 1
    assert lib.abrb_frame_id(h, b"EE") == 13 and lib.abrb_frame_id(h, b"joint2") == 9
    assert lib.abrb_frame_id(h, b"link7") == _abi.EFRAME
    assert b"Invalid transformation name" in lib.abrb_last_error()
    # bad descriptors
    bad = _abi.chain_desc_from_dict(_abi.load_arm_json("ur5"))
    bad.n_links = 6
    h2 = C.c_void_p()
    assert lib.abrb_model_create(C.byref(bad), C.byref(h2)) == _abi.ESHAPE and not h2
    bad.n_joints = 9
    assert lib.abrb_model_create(C.byref(bad), C.byref(h2)) == _abi.ESHAPE
    # controller parameter checks mirror the reference's exceptions
    hc = C.c_void_p()
    p = _abi.osc_params(6, kp=10, orientation_algorithm=3)
    assert lib.abrb_osc_create(h, C.byref(p), C.byref(hc)) == _abi.EUNSUP
    assert b"Invalid algorithm number" in lib.abrb_last_error()
    p = _abi.osc_params(6, kp=10)
    assert lib.abrb_osc_create(h, C.byref(p), C.byref(hc)) == 0
    # argument validation happens before any device work
    out = _abi.RbdOut()
    assert lib.abrb_rbd_eval_f64(h, 99, None, None, None, 4, C.byref(out), None) == _abi.EFRAME
    assert lib.abrb_rbd_eval_f64(h, 13, None, None, None, -1, C.byref(out), None) == _abi.EINVAL
    assert lib.abrb_rbd_eval_f64(h, 13, None, None, None, 0, C.byref(out), None) == 0  # empty batch is a no-op
    assert lib.abrb_osc_generate_f64(hc, 13, None, None, None, None, 6, None, 0, None, None, None, 0, None) == 0
    assert lib.abrb_osc_generate_f64(hc, 13, None, None, None, None, 5, None, 0, None, None, None, 8, None) == _abi.EINVAL
    # integrated_error goes with ki != 0 and only with it (osc.py:81-82, :262-264)
    dummy = (C.c_double * 48)()
    assert lib.abrb_osc_generate_f64(hc, 13, None, None, None, None, 6, None, 0, None, None, dummy, 8, None) == _abi.EINVAL
    assert b"integrated_error" in lib.abrb_last_error()
    pk = _abi.osc_params(6, kp=10, ki=0.5)
    hk = C.c_void_p()
    assert lib.abrb_osc_create(h, C.byref(pk), C.byref(hk)) == 0
    assert lib.abrb_osc_generate_f64(hk, 13, None, None, None, None, 6, None, 0, None, None, None, 8, None) == _abi.EINVAL
    assert lib.abrb_osc_rollout_f64(hk, 13, None, None, None, None, 6, 4, 1e-3, None, None, None, None, 8, None) == _abi.EINVAL
    assert lib.abrb_osc_generate_host_async_f64(hc, 13, None, None, None, None, 6, None, 0, None, None, None, 8, 2) == _abi.EINVAL
    assert lib.abrb_osc_host_wait(hc, 0) == 0 and lib.abrb_osc_host_wait(hc, 5) == _abi.EINVAL
    assert lib.abrb_osc_set_option(hc, b"host_chunk_states", 32768.0) == 0
    for k in (1.0, 2.0, 3.0):
        assert lib.abrb_osc_set_option(hc, b"host_upload_streams", k) == 0
    for k in (0.0, 4.0, float("nan")):
        assert lib.abrb_osc_set_option(hc, b"host_upload_streams", k) == _abi.EINVAL
        assert b"host_upload_streams" in lib.abrb_last_error()
    assert lib.abrb_osc_set_option(hc, b"no_such_option", 1.0) == _abi.EINVAL
    assert lib.abrb_osc_destroy(hk) == 0
    assert lib.abrb_osc_destroy(hc) == 0 and lib.abrb_model_destroy(h) == 0
    jaco = _abi.chain_desc_from_dict(_abi.load_arm_json("jaco2"))
    assert lib.abrb_model_create(C.byref(jaco), C.byref(h)) == 0
    assert lib.abrb_model_is_orthonormal(h) == 0  # measured frames
    lib.abrb_model_destroy(h)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly (ABRB_ECUDA), never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from abr_control_b200.arms import ur5
    from abr_control_b200.controllers import OSC

    rc = ur5.Config()
    with pytest.raises(_lib.AbrbError) as ei:
        rc.M(np.zeros(6))
    assert ei.value.code == _abi.ECUDA
    with pytest.raises(_lib.AbrbError):
        OSC(rc, kp=10).generate(np.zeros(6), np.zeros(6), np.zeros(6))


def test_package_does_not_import_the_oracle():
    import sys

    import abr_control_b200  # noqa: F401
    for root, _, files in os.walk(os.path.join(ROOT, "abr_control_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/ may", ""), f
