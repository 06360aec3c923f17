"""Chain descriptor from a MuJoCo MJCF file.

The reference's second arm-model implementation, ``MujocoConfig`` (/root/reference/abr_control/arms/mujoco_config.py),
reads an MJCF file (:64-117: ``custom/numeric`` START_ANGLES, the ``actuator`` list naming the arm's joints) and then asks
the MuJoCo simulator for ``J``, ``M``, ``g`` ... at run time.  Here the same file is turned, once, into the flat chain
descriptor of include/abrb.h (``abrb_chain_desc``), so that the batched kernels evaluate the arm without a simulator:

    link0 = L0,   joint_i = link_i . A[i],   link_{i+1} = joint_i . Rz(q_i) . B[i],   EE = link_n . E

* the arm is the serial chain of ``<body>`` elements that carry the actuated hinge joints, in actuator order;
* joint frame i: origin at the joint's ``pos`` in its body, z axis along the joint's ``axis``;
* link frame i+1: the body's ``<inertial>`` frame (``pos``, optional ``quat``); ``mass`` and ``diaginertia`` fill
  ``link_inertia`` (``diag(m, m, m, Ixx, Iyy, Izz)``, the layout of the reference's ``_M_LINKS``, base_config.py:625-632);
* EE: the body named ``ee_body`` (default "EE") below the last link, else the last link's body frame.

Like the reference's SymPy configs (and unlike a simulator) the engine does not rotate the diagonal inertia
(SURVEY.md S0.6), so ``M``, ``g``, ``C`` agree with MuJoCo's for isotropic ``diaginertia`` (all shipped UR5 links:
0.1 0.1 0.1) and with the reference's formula otherwise.  MuJoCo itself is absent from this image: the importer is
pinned through the kinematics it must share with the reference's SymPy config of the same arm (tests/test_mjcf.py).
Supported: ``<compiler angle="radian"|"degree">``, body ``pos`` / ``quat`` / ``euler`` (MuJoCo's default xyz sequence),
hinge joints; not supported (ValueError): slide / ball / free joints inside the arm, branching arms, ``<include>``.
"""
import xml.etree.ElementTree as ElementTree

import numpy as np


def _floats(text, n, default):
    if text is None:
        return np.array(default, dtype=float)
    v = np.array([float(x) for x in text.split()], dtype=float)
    if v.size != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return v


def _quat_to_R(qw):
    w, x, y, z = qw / np.linalg.norm(qw)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _euler_to_R(e):
    """MuJoCo's default eulerseq "xyz": intrinsic rotations about x, then y, then z"""
    cx, sx, cy, sy, cz, sz = np.cos(e[0]), np.sin(e[0]), np.cos(e[1]), np.sin(e[1]), np.cos(e[2]), np.sin(e[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def _frame(elem, degrees):
    """4x4 transform of an element's pos / quat / euler attributes relative to its parent"""
    T = np.eye(4)
    T[:3, 3] = _floats(elem.get("pos"), 3, [0, 0, 0])
    if elem.get("quat") is not None:
        T[:3, :3] = _quat_to_R(_floats(elem.get("quat"), 4, [1, 0, 0, 0]))
    elif elem.get("euler") is not None:
        e = _floats(elem.get("euler"), 3, [0, 0, 0])
        T[:3, :3] = _euler_to_R(np.deg2rad(e) if degrees else e)
    elif any(elem.get(k) is not None for k in ("axisangle", "xyaxes", "zaxis")):
        raise ValueError("orientation given as axisangle / xyaxes / zaxis is not supported")
    return T


def _z_to(axis):
    """rotation whose third column is the unit vector `axis` (the joint frame's z axis)"""
    z = axis / np.linalg.norm(axis)
    helper = np.array([1.0, 0, 0]) if abs(z[0]) < 0.9 else np.array([0, 1.0, 0])
    x = np.cross(helper, z)
    x /= np.linalg.norm(x)
    return np.column_stack([x, np.cross(z, x), z])


def _inv(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def chain_desc_from_mjcf(xml_file, ee_body="EE", gravity=(0.0, 0.0, -9.81)):
    """-> dict accepted by ``BaseConfig`` / ``_abi.chain_desc_from_dict`` (plus ``start_angles``, ``joint_names``)"""
    root = ElementTree.parse(xml_file).getroot()
    compiler = root.find("compiler")
    degrees = compiler is None or compiler.get("angle", "degree") == "degree"
    actuators = root.find("actuator")
    if actuators is None:
        raise ValueError("the MJCF file names no actuators: which joints form the arm?")
    joint_names = [a.get("joint") for a in actuators]
    start = None
    n_grip = 0
    for num in root.findall("custom/numeric"):
        if num.get("name") == "START_ANGLES":
            start = [float(v) for v in num.get("data").split()]
        elif num.get("name") == "N_GRIPPER_JOINTS":
            n_grip = int(float(num.get("data")))
    if n_grip:
        joint_names = joint_names[: len(joint_names) - n_grip]  # mujoco_config.py keeps the gripper out of the arm
    n = len(joint_names)

    # walk the body tree keeping the world transform of every body (all joints at zero)
    chain = []  # (world transform of the body, <body>, <joint>) in the order the joints are met

    def walk(body, T_parent):
        T = T_parent @ _frame(body, degrees)
        for j in body.findall("joint"):
            if j.get("name") in joint_names:
                if j.get("type", "hinge") != "hinge":
                    raise ValueError(f"joint {j.get('name')}: only hinge joints are supported")
                chain.append((T, body, j))
        for child in body.findall("body"):
            walk(child, T)

    world = root.find("worldbody")
    for b in world.findall("body"):
        walk(b, np.eye(4))
    if [j.get("name") for _, _, j in chain] != joint_names:
        raise ValueError("the actuated joints do not form one serial chain in actuator order")
    for (_, ba, _), (_, bb, _) in zip(chain[:-1], chain[1:]):
        if bb not in list(ba.iter("body")):
            raise ValueError("branching arms are not supported")

    def parent_of(body):
        for cand in world.iter("body"):
            if body in cand.findall("body"):
                return cand
        return None

    # link0: the body that carries the first joint's body (the base), at its body frame; its inertia does not enter M
    # (link 0 moves with no joint) but its row is kept for the reference's N_LINKS = n + 1 layout
    base = parent_of(chain[0][1])
    T_base = np.eye(4)
    if base is not None:
        anc, node = [], base
        while node is not None:
            anc.append(node)
            node = parent_of(node)
        for b in reversed(anc):
            T_base = T_base @ _frame(b, degrees)
    link_T = [T_base]      # world transform of link frames (COM frames), joints at zero
    joint_T = []           # world transform of joint frames (z = axis), joints at zero
    inertia = [[0.0] * 6]
    if base is not None and base.find("inertial") is not None:
        m = float(base.find("inertial").get("mass", 0))
        di = _floats(base.find("inertial").get("diaginertia"), 3, [0, 0, 0])
        inertia[0] = [m, m, m, di[0], di[1], di[2]]
    for T_body, body, j in chain:
        Tj = np.eye(4)
        Tj[:3, :3] = _z_to(_floats(j.get("axis"), 3, [0, 0, 1]))
        Tj[:3, 3] = _floats(j.get("pos"), 3, [0, 0, 0])
        joint_T.append(T_body @ Tj)
        inert = body.find("inertial")
        if inert is None:
            raise ValueError(f"body {body.get('name')} has no <inertial> (mesh-derived inertias need the simulator)")
        if inert.get("fullinertia") is not None:
            raise ValueError("fullinertia is not supported (give diaginertia and the inertial frame's quat)")
        link_T.append(T_body @ _frame(inert, degrees))
        m = float(inert.get("mass"))
        di = _floats(inert.get("diaginertia"), 3, [0, 0, 0])
        inertia.append([m, m, m, di[0], di[1], di[2]])
    T_ee = chain[-1][0]
    for b in chain[-1][1].iter("body"):
        if b.get("name") == ee_body:
            T_ee, node, path = np.eye(4), b, []
            while node is not None and node is not chain[-1][1]:
                path.append(node)
                node = parent_of(node)
            T_ee = chain[-1][0]
            for p in reversed(path):
                T_ee = T_ee @ _frame(p, degrees)
    blk = lambda T: [[float(v) for v in row] for row in T[:3, :4]]  # noqa: E731
    # with every joint at zero:  joint_i = link_i . A[i]  and  link_{i+1} = joint_i . B[i]
    A = [blk(_inv(link_T[i]) @ joint_T[i]) for i in range(n)]
    B = [blk(_inv(joint_T[i]) @ link_T[i + 1]) for i in range(n)]
    return dict(name=root.get("model", "mjcf"), n_joints=n, n_links=n + 1, L0=blk(link_T[0]), A=A, B=B,
                E=blk(_inv(link_T[n]) @ T_ee), link_inertia=inertia,
                gravity=[float(gravity[0]), float(gravity[1]), float(gravity[2]), 0.0, 0.0, 0.0],
                start_angles=start if start is not None else [0.0] * n, joint_names=joint_names)


class MjcfConfig:
    """``robot_config`` built from an MJCF file: ``MjcfConfig("ur5.xml")`` -> a batched ``BaseConfig``"""

    def __new__(cls, xml_file, ee_body="EE", dtype=np.float64, **kw):
        from .base_config import BaseConfig

        desc = chain_desc_from_mjcf(xml_file, ee_body=ee_body)
        return BaseConfig(desc, ROBOT_NAME=desc["name"], dtype=dtype, **kw)
