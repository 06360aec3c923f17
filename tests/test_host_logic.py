"""CPU suite: host-side logic of the Python mirrors that needs no device — constructor semantics that must match the
reference's (checked against the oracle restatements, which are pinned to reference goldens), parameter marshalling into
the C structs, option handling.  Model / controller handles can be created without a GPU (no compute happens)."""
import ctypes as C

import numpy as np
import pytest

import cases
from abr_control_b200 import _abi, _lib
from abr_control_b200.arms import jaco2, threejoint, ur5
from abr_control_b200.controllers import OSC, AvoidJointLimits, Damping, Sliding
from abr_control_b200.controllers.path_planners import InverseKinematics
from oracle import osc_oracle as oo

ARM = {"ur5": ur5, "jaco2": jaco2, "threejoint": threejoint}


@pytest.mark.parametrize("name", [k for k, v in cases.NULL_CASES.items() if v["ctrl"][0] == "AvoidJointLimits"])
def test_avoid_joint_limits_constructor_matches_the_reference_semantics(name):
    """limits shifted by -pi, swapped where cross_zero, NaN = no limit (avoid_joint_limits.py:36-86)"""
    cs = cases.NULL_CASES[name]
    kind, kw = cs["ctrl"]
    rc = ARM[cs["arm"]].Config()
    mine = AvoidJointLimits(rc, **kw)
    ref = oo.AvoidJointLimits(oo.RobotOracle(cs["arm"]), **kw)
    np.testing.assert_array_equal(mine.min_joint_angles, ref.lo)
    np.testing.assert_array_equal(mine.max_joint_angles, ref.hi)
    np.testing.assert_array_equal(mine.no_limits_min, ref.no_lo)
    np.testing.assert_array_equal(mine.no_limits_max, ref.no_hi)
    np.testing.assert_array_equal(mine.max_torque, ref.tmax)
    p = mine._params()
    n = rc.N_JOINTS
    assert p.kind == _abi.NULL_JOINT_LIMITS
    assert [bool(p.limit_cross_zero[k]) for k in range(n)] == list(map(bool, kw.get("cross_zero", [False] * n)))
    assert [bool(p.limit_gradient[k]) for k in range(n)] == list(map(bool, kw.get("gradient", [False] * n)))


def test_avoid_joint_limits_rejects_wrong_sizes_and_accepts_none():
    rc = ur5.Config()
    with pytest.raises(Exception, match="joint angles vector incorrect size"):
        AvoidJointLimits(rc, [0.1] * 5, [1.0] * 6)
    a = AvoidJointLimits(rc, [None, 0.5, None, 1.0, None, 2.0], [3.0, None, 4.0, None, 5.0, None])
    assert list(a.no_limits_min) == [True, False, True, False, True, False]
    assert list(a.no_limits_max) == [False, True, False, True, False, True]
    assert np.allclose(a.max_torque, 1.0)


def test_osc_options_survive_parameter_changes():
    rc = ur5.Config()
    damp = Damping(rc, kv=10)
    c = OSC(rc, kp=10, ctrlr_dof=[True] * 6, null_controllers=[damp])
    c.set_option("host_chunk_states", 32768)
    h1 = c._native()
    assert h1 and c._options == {"host_chunk_states": 32768.0}
    damp.kv = 5.0  # a secondary controller changed: the owner drops its native handle ...
    assert c._handle is None
    h2 = c._native()
    assert h2 and c._options == {"host_chunk_states": 32768.0}  # ... rebuilds it and re-applies its options
    c.use_g = False  # the reference reads the controller's own attributes per call too (osc.py:300)
    assert c._handle is None and c._native()
    c.record_training_signal = False  # not a parameter of the native handle
    assert c._handle is not None
    with pytest.raises(RuntimeError):
        c.set_option("no_such_option", 1)


def test_ki_keeps_the_reference_attribute_and_owners_are_weak():
    import gc
    import weakref

    rc = ur5.Config()
    damp = Damping(rc, kv=10)
    c = OSC(rc, kp=10, ki=0.1, null_controllers=[damp])
    assert c.integrated_error.shape == (6,) and not c.integrated_error.any()  # osc.py:81-82
    assert c._native()
    r = weakref.ref(c)
    del c
    gc.collect()
    assert r() is None and len(damp._owners) == 0  # a secondary controller does not keep its OSCs alive


def test_osc_parameter_marshalling_of_joint_limits_inside_osc():
    cs = cases.OSC_CASES["ur5_limits_grad"]
    rc = ur5.Config()
    nulls = [AvoidJointLimits(rc, **kw) for _, kw in cs["null"]]
    c = OSC(rc, null_controllers=nulls, **cs["osc"])
    assert c._native()  # the library accepts the embedded secondary controller
    p = _abi.osc_params(6, null=[nc._params() for nc in nulls], **cs["osc"])
    assert p.n_null == 1 and p.null[0].kind == _abi.NULL_JOINT_LIMITS
    assert np.isnan(p.null[0].limit_min[1]) and np.isnan(p.null[0].limit_max[2])
    assert p.null[0].limit_min[3] == pytest.approx(1.0 - np.pi) and p.null[0].limit_max[3] == pytest.approx(5.5 - np.pi)  # swapped


def test_planner_and_sliding_argument_checks_need_no_device():
    rc = ur5.Config()
    ik = InverseKinematics(rc, max_dx=0.3)
    assert (ik.max_dx, ik.max_dr, ik.max_dq) == (0.3, 2 * np.pi, np.pi)
    with pytest.raises(ValueError, match="method must be 1, 2 or 3"):
        ik.generate_path(np.zeros(6), np.zeros(6), method=4)
    s = Sliding(rc)
    assert (s.kd, s.lamb, s.cartesian) == (160.0, 30.0, True)
    # C ABI validation happens before any device work
    L = _lib.lib()
    assert L.abrb_ik_path_f64(rc.handle, 0.2, 6.28, 3.14, 7, 0.001, 10, None, None, 6, None, None, 4, None) == _abi.EUNSUP
    assert L.abrb_ik_path_f64(rc.handle, 0.2, 6.28, 3.14, 3, 0.001, 10, None, None, 5, None, None, 4, None) == _abi.EINVAL
    assert L.abrb_ik_path_f64(rc.handle, 0.2, 6.28, 3.14, 3, 0.001, 0, None, None, 6, None, None, 4, None) == 0
    assert L.abrb_sliding_generate_f64(rc.handle, 160.0, 30.0, 1, 99, None, None, None, None, 3, None, 0, None, 0, None,
                                       None, 4, None) == _abi.EFRAME
    assert L.abrb_sliding_generate_f64(rc.handle, 160.0, 30.0, 1, 13, None, None, None, None, 6, None, 0, None, 0, None,
                                       None, 4, None) == _abi.EINVAL  # cartesian rows have 3 values
    assert L.abrb_sliding_generate_f64(rc.handle, 160.0, 30.0, 0, 13, None, None, None, None, 6, None, 0, None, 0, None,
                                       None, 0, None) == 0  # empty batch
