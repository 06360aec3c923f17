"""Shared definition of the parity cases.

The same table drives (i) ``oracle/ref_harness/run_reference.py`` — which builds the REFERENCE's
controllers from it inside the development container and writes ``tests/golden/*.npz`` — and
(ii) the parity tests, which build this repo's controllers from it.  Pure data: no imports from
either implementation.

Controller parameterisations follow the reference's examples:
  * config 1  /root/reference/examples/PyGame/force_osc_xy.py:20-34
  * config 3  /root/reference/examples/timing_plots.py:37 + examples/CoppeliaSim/force_osc_xyzabg.py:16-26
  * config 5  /root/reference/examples/CoppeliaSim/force_osc_xyz_avoid_obstacle.py:17-28,40,74
"""
import numpy as np

N_GOLDEN = 32  # states per golden file (kept small: fixtures are committed)
XOFF = [0.11, -0.07, 0.05]  # non-zero offset inside a frame, exercises the general-x functions

ARMS = {
    "ur5": dict(n=6, has_C=True, mid="link3"),
    "jaco2": dict(n=6, has_C=False, mid="link4"),  # C oracle un-generatable
    "threejoint": dict(n=3, has_C=True, mid="link2"),
    "twojoint": dict(n=2, has_C=True, mid="link1"),
}


def frames(n):
    return [f"link{i}" for i in range(n + 1)] + [f"joint{i}" for i in range(n)] + ["EE"]


def states(arm, count=N_GOLDEN):
    """Seeded synthetic joint states: q~U(0,2pi), dq~U(-5,5), target~U(-1,1) (cf. timing_plots.py:18-20)."""
    n = ARMS[arm]["n"]
    rng = np.random.default_rng(abs(hash_name(arm)))
    q = rng.uniform(0, 2 * np.pi, (count, n))
    dq = rng.uniform(-5, 5, (count, n))
    target = rng.uniform(-1, 1, (count, 6))
    target_velocity = rng.uniform(-0.5, 0.5, (count, 6))
    return q, dq, target, target_velocity


def hash_name(s):
    h = 0
    for ch in s:
        h = (h * 131 + ord(ch)) % 1000003
    return h


T, F = True, False
PI = float(np.pi)
NAN = float("nan")  # AvoidJointLimits: no limit on that side (the reference's None placeholder fails its own isnan)

# name -> dict(arm, osc=<OSC kwargs>, null=[(kind, kwargs)...], tv=<use target_velocity>,
#              ref_frame, xyz_offset)
OSC_CASES = {
    # --- UR5 -----------------------------------------------------------------------------------
    "ur5_xyz": dict(arm="ur5", osc=dict(kp=10)),
    "ur5_6dof": dict(arm="ur5", osc=dict(kp=30, ko=20, ctrlr_dof=[T] * 6)),
    "ur5_6dof_C_damp": dict(
        arm="ur5", osc=dict(kp=50, ctrlr_dof=[T] * 6, use_C=True), null=[("Damping", dict(kv=10))]
    ),
    "ur5_6dof_alg1": dict(arm="ur5", osc=dict(kp=25, kv=9, ctrlr_dof=[T] * 6, orientation_algorithm=1)),
    "ur5_vmax": dict(arm="ur5", osc=dict(kp=10, ko=8, kv=4, ctrlr_dof=[T] * 6, vmax=[0.5, 1.0])),
    "ur5_tv_offset": dict(
        arm="ur5",
        osc=dict(kp=20, ctrlr_dof=[T, T, T, F, T, F]),
        tv=True,
        ref_frame="link4",
        xyz_offset=XOFF,
    ),
    "ur5_nog_xz": dict(arm="ur5", osc=dict(kp=5, kv=3, ctrlr_dof=[T, F, T, F, F, F], use_g=False)),
    "ur5_rest": dict(
        arm="ur5",
        osc=dict(kp=40),
        null=[("RestingConfig", dict(kp=30, kv=6, rest_angles=[None, PI / 4, -PI / 2, PI / 4, None, None]))],
    ),
    # --- Jaco2 ---------------------------------------------------------------------------------
    "jaco2_cfg3": dict(
        arm="jaco2", osc=dict(kp=200, ctrlr_dof=[T] * 5 + [F]), null=[("Damping", dict(kv=10))]
    ),
    "jaco2_cfg5": dict(
        arm="jaco2",
        osc=dict(kp=200, vmax=[0.5, 0], ctrlr_dof=[T, T, T, F, F, F]),
        null=[
            ("AvoidObstacles", dict(obstacles=[[0.09596, -0.2661, 0.64204, 0.05]], threshold=0.2)),
            ("Damping", dict(kv=10)),
        ],
    ),
    "jaco2_obst2": dict(
        arm="jaco2",
        osc=dict(kp=100, ctrlr_dof=[T, T, T, F, F, F]),
        null=[
            (
                "AvoidObstacles",
                dict(
                    obstacles=[[0.1, -0.2, 0.5, 0.08], [-0.15, 0.1, 0.6, 0.1]],
                    threshold=0.5,
                    gain=2.0,
                    maximum=40.0,
                ),
            )
        ],
    ),
    "jaco2_alg1_tv": dict(
        arm="jaco2", osc=dict(kp=60, ko=40, kv=12, ctrlr_dof=[T] * 6, orientation_algorithm=1), tv=True
    ),
    # --- threejoint (config 1) and twojoint (the arm the reference's unit tests pin) ----------------
    "threejoint_cfg1": dict(
        arm="threejoint",
        osc=dict(kp=20, use_C=True, ctrlr_dof=[T, T, F, F, F, F]),
        null=[
            ("Damping", dict(kv=10)),
            ("RestingConfig", dict(kp=50, kv=float(np.sqrt(50)), rest_angles=[PI / 4, PI, None])),
        ],
    ),
    "twojoint_xy": dict(arm="twojoint", osc=dict(kp=15, use_C=True, ctrlr_dof=[T, T, F, F, F, F])),
    # examples/PyGame/force_osc_xy_avoid_joint_limits.py:21-38
    "threejoint_limits": dict(
        arm="threejoint",
        osc=dict(kp=100, ctrlr_dof=[T, T, F, F, F, F]),
        null=[
            ("AvoidJointLimits", dict(min_joint_angles=[PI / 5.0] * 3, max_joint_angles=[PI / 2.0] * 3, max_torque=[100.0] * 3)),
            ("Damping", dict(kv=10)),
        ],
    ),
    "ur5_limits_grad": dict(
        arm="ur5",
        osc=dict(kp=40, ko=30, ctrlr_dof=[T] * 6),
        null=[("AvoidJointLimits", dict(
            min_joint_angles=[0.5, NAN, 1.0, 5.5, NAN, 2.0], max_joint_angles=[2.0, 3.0, NAN, 1.0, NAN, 4.0],
            max_torque=[3.0, 4.0, 5.0, 6.0, 7.0, 8.0], cross_zero=[F, F, F, T, F, F], gradient=[T, F, F, T, F, T]))],
    ),
}

# standalone secondary controllers
NULL_CASES = {
    "ur5_damping": dict(arm="ur5", ctrl=("Damping", dict(kv=10))),
    "ur5_resting": dict(
        arm="ur5", ctrl=("RestingConfig", dict(kp=30, kv=6, rest_angles=[None, PI / 4, -PI / 2, PI / 4, None, None]))
    ),
    "jaco2_avoid": dict(
        arm="jaco2",
        ctrl=("AvoidObstacles", dict(obstacles=[[0.1, -0.2, 0.5, 0.08], [-0.15, 0.1, 0.6, 0.1]], threshold=0.5)),
    ),
    "ur5_avoid": dict(
        arm="ur5", ctrl=("AvoidObstacles", dict(obstacles=[[0.2, 0.1, 0.4, 0.05]], threshold=0.6, gain=1.5, maximum=30.0))
    ),
    "threejoint_resting": dict(
        arm="threejoint", ctrl=("RestingConfig", dict(kp=50, kv=float(np.sqrt(50)), rest_angles=[PI / 4, PI, None]))
    ),
    "ur5_limits": dict(arm="ur5", ctrl=("AvoidJointLimits", dict(
        min_joint_angles=[0.5, NAN, 1.0, 5.5, NAN, 2.0], max_joint_angles=[2.0, 3.0, NAN, 1.0, NAN, 4.0],
        max_torque=[3.0, 4.0, 5.0, 6.0, 7.0, 8.0], cross_zero=[F, F, F, T, F, F], gradient=[T, F, F, T, F, T]))),
    "jaco2_limits_wall": dict(arm="jaco2", ctrl=("AvoidJointLimits", dict(
        min_joint_angles=[1.0, 0.8, 0.3, NAN, 5.0, 1.5], max_joint_angles=[5.0, 5.4, 6.0, NAN, 1.2, 4.5],
        cross_zero=[F, F, F, F, T, F]))),
}

# joint-space controllers: Joint.generate(q, dq, target, target_velocity) and Floating.generate(q, dq)
CTRL_CASES = {
    "ur5_joint": dict(arm="ur5", ctrl=("Joint", dict(kp=25, kv=7)), tv=True),
    "jaco2_joint_nograv": dict(arm="jaco2", ctrl=("Joint", dict(kp=10, account_for_gravity=False))),
    "ur5_floating": dict(arm="ur5", ctrl=("Floating", dict())),
    "ur5_floating_task_dyn": dict(arm="ur5", ctrl=("Floating", dict(task_space=True, dynamic=True))),
    "jaco2_floating_task": dict(arm="jaco2", ctrl=("Floating", dict(task_space=True))),
    "threejoint_joint": dict(arm="threejoint", ctrl=("Joint", dict(kp=50))),
}


# Sliding.generate(q, dq, target, target_velocity, target_acc, ref_frame, offset)  
#   cartesian: target/velocity/acc are the first 3 columns of states()'s target / target_velocity / joint_targets()[1];
#   joint space: joint_targets() and dq-like rows
SLIDING_CASES = {
    "ur5_sliding_xyz": dict(arm="ur5", ctrl=dict(kd=10.0, lamb=30.0)),
    "ur5_sliding_xyz_full": dict(arm="ur5", ctrl=dict(kd=160.0, lamb=30.0), tv=True, ta=True, ref_frame="EE", offset=XOFF),
    "ur5_sliding_joint": dict(arm="ur5", ctrl=dict(kd=20.0, lamb=5.0, cartesian=False), tv=True, ta=True),
    "threejoint_sliding_xyz": dict(arm="threejoint", ctrl=dict(kd=160.0, lamb=30.0), tv=True),
}


def sliding_inputs(case, count=N_GOLDEN):
    """-> (target, target_velocity or None, target_acc or None) rows for one SLIDING_CASES entry"""
    arm = case["arm"]
    _, _, target, tvel = states(arm, count)
    tq, tdq = joint_targets(arm, count)
    if case["ctrl"].get("cartesian", True):
        tgt, tv, ta = target[:, :3], tvel[:, :3], 0.3 * tdq[:, :1] + tvel[:, 3:6]
    else:
        tgt, tv, ta = tq, tdq, 0.5 * tdq[:, ::-1]
    return (np.ascontiguousarray(tgt), np.ascontiguousarray(tv) if case.get("tv") else None,
            np.ascontiguousarray(ta) if case.get("ta") else None)


# InverseKinematics(robot_config, **init).generate_path(position, target_position, **path)  (SURVEY.md S8f#3);
# positions = states()[0], targets = ik_targets(); short horizons keep the fixtures small
IK_CASES = {
    "ur5_ik_m3": dict(arm="ur5", path=dict(n_timesteps=24, dt=0.05, method=3)),
    "ur5_ik_m2": dict(arm="ur5", path=dict(n_timesteps=24, dt=0.05, method=2)),
    "ur5_ik_m1": dict(arm="ur5", path=dict(n_timesteps=24, dt=0.02, method=1)),
    "ur5_ik_default_dt": dict(arm="ur5", init=dict(max_dx=0.5, max_dr=1.0, max_dq=2.0), path=dict(n_timesteps=16, method=3)),
    "jaco2_ik_m3": dict(arm="jaco2", path=dict(n_timesteps=24, dt=0.05, method=3)),
}
N_IK = 12  # trajectories per IK case


def ik_targets(arm, count=N_IK):
    """task-space targets inside the arm's reach: xyz in a box, Euler angles in (-pi, pi)"""
    rng = np.random.default_rng(hash_name(arm) + 29)
    xyz = rng.uniform(-0.45, 0.45, (count, 3)) + np.array([0.0, 0.0, 0.45])
    return np.hstack([xyz, rng.uniform(-np.pi, np.pi, (count, 3))])


def joint_targets(arm, count=N_GOLDEN):
    n = ARMS[arm]["n"]
    rng = np.random.default_rng(hash_name(arm) + 17)
    return rng.uniform(-2 * np.pi, 4 * np.pi, (count, n)), rng.uniform(-1, 1, (count, n))
