"""CPU: batched kinematics + the end-effector controllers (mani_skill/agents/controllers/pd_ee_pose.py, utils/kinematics.py GPU
branch).  pytorch_kinematics is absent, so the chain is checked against the simulator's own forward kinematics (link poses fetched
from the backend), the Jacobian against finite differences, and the controllers by what they are for: the end effector moves by
the commanded displacement."""
import numpy as np
import pytest
import torch

import maniskill_b200 as ms
from maniskill_b200 import utils as U
from maniskill_b200.kinematics import Kinematics, SerialChain
from maniskill_b200.model import load_robot
from maniskill_b200.structs import Pose

from emu_world import EmuBackendWorld


def _env(mode, n=4):
    return ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode=mode, device="cpu", world_factory=EmuBackendWorld)


def test_chain_fk_matches_the_simulator_and_jacobian_matches_finite_differences():
    env = _env("pd_joint_delta_pos")
    env.reset(seed=1)
    robot = env.agent.robot
    kin = Kinematics(load_robot("panda_v2"), "panda_hand_tcp", robot.dof_names, env.agent.arm_joint_names, env.device)
    assert kin.chain.joint_names == env.agent.arm_joint_names
    q = robot.get_qpos()
    p, r = kin.fk(q)
    ee_at_base = robot.root.pose.inv() * env.agent.tcp.pose
    assert torch.allclose(p, ee_at_base.p, atol=2e-6)
    same = torch.minimum((r - ee_at_base.q).abs().max(dim=1)[0], (r + ee_at_base.q).abs().max(dim=1)[0])  # q and -q: same rotation
    assert same.max() < 2e-6
    # geometric Jacobian (linear rows, then angular) against central differences of the chain itself, in float64
    chain = SerialChain(load_robot("panda_v2"), "panda_hand_tcp", "cpu", torch.float64)
    q64 = q[:, kin.chain_dof_idx].double()
    p0, R0, J = chain.forward(q64)
    eps = 1e-6
    for j in range(chain.n_joints):
        dq = torch.zeros_like(q64)
        dq[:, j] = eps
        pp, Rp, _ = chain.forward(q64 + dq)
        pm, Rm, _ = chain.forward(q64 - dq)
        lin = (pp - pm) / (2 * eps)
        dR = Rp @ Rm.transpose(1, 2)                  # rotation between the two perturbed frames, root frame
        ang = torch.stack([dR[:, 2, 1] - dR[:, 1, 2], dR[:, 0, 2] - dR[:, 2, 0], dR[:, 1, 0] - dR[:, 0, 1]], 1) / 2 / (2 * eps)  # small angle
        assert torch.allclose(J[:, :3, j], lin, atol=1e-6), j
        assert torch.allclose(J[:, 3:, j], ang, atol=1e-6), j


def test_ik_step_reaches_a_small_displacement():
    env = _env("pd_joint_delta_pos")
    env.reset(seed=2)
    robot = env.agent.robot
    kin = Kinematics(load_robot("panda_v2"), "panda_hand_tcp", robot.dof_names, env.agent.arm_joint_names, env.device)
    q = robot.get_qpos()
    p0, r0 = kin.fk(q)
    delta = torch.tensor([[0.02, -0.01, 0.015, 0.0, 0.0, 0.0]]).expand(q.shape[0], 6)
    for solver in ("levenberg_marquardt", "pseudo_inverse"):
        tq = kin.compute_ik(delta, q, dict(type=solver, alpha=1.0))
        q2 = q.clone()
        q2[:, :7] = tq
        p1, r1 = kin.fk(q2)
        assert torch.allclose(p1 - p0, delta[:, :3], atol=2e-3), solver      # first-order step: error is second order in |delta|
        assert (U.quat_mul(r1, U.quat_conj(r0))[:, 1:]).abs().max() < 2e-3, solver


@pytest.mark.parametrize("mode", ["pd_ee_delta_pos", "pd_ee_target_delta_pos", "pd_ee_delta_pose", "pd_ee_target_delta_pose"])
def test_ee_controllers_move_the_tcp_by_the_commanded_displacement(mode):
    env = _env(mode)
    env.reset(seed=3)
    n = env.num_envs
    assert env.action_dim == (4 if mode.endswith("pos") else 7)
    tcp0 = env.agent.tcp.pose.raw_pose.clone()
    a = torch.zeros(n, env.action_dim)
    a[:, 0], a[:, 2] = 0.5, -0.4          # +5 cm along x, -4 cm along z of the root frame per control step (bounds +-0.1)
    a[:, -1] = 1.0                        # keep the gripper open
    for _ in range(3):
        obs, rew, term, trunc, info = env.step(a)
    want = torch.tensor([0.15, 0.0, -0.12])
    if "target" in mode:
        # the virtual target accumulates the commands; give the (overdamped, kd/kp = 0.1 s) PD drive time to reach it
        hold = torch.zeros(n, env.action_dim)
        hold[:, -1] = 1.0
        for _ in range(12):
            env.step(hold)
        moved = env.agent.tcp.pose.p - tcp0[:, :3]
        assert (moved - want).abs().max() < 0.01, moved
    else:
        # every command is relative to the current pose: the drive lag is not compensated, the tcp moves along the commanded
        # direction and covers a good part of it
        moved = env.agent.tcp.pose.p - tcp0[:, :3]
        cosang = (moved * want).sum(1) / (moved.norm(dim=1) * want.norm())
        assert cosang.min() > 0.995, cosang
        assert 0.25 < (moved.norm(dim=1) / want.norm()).min() and (moved.norm(dim=1) / want.norm()).max() < 1.05
    # orientation is held
    dq = U.quat_mul(env.agent.tcp.pose.q, U.quat_conj(tcp0[:, 3:]))
    assert dq[:, 1:].abs().max() < 0.03
    if mode.endswith("pose"):
        # rotation command about the root z axis: the reference scales the (norm-clipped) rotation action by rot_lower = -0.1
        # (pd_ee_pose.py:229-239), i.e. +1 commands -0.1 rad per step
        q_before = env.agent.tcp.pose.q.clone()
        a2 = torch.zeros(n, 7)
        a2[:, 5], a2[:, -1] = 1.0, 1.0
        for _ in range(3):
            env.step(a2)
        if "target" in mode:
            hold = torch.zeros(n, 7)
            hold[:, -1] = 1.0
            for _ in range(12):
                env.step(hold)
        d = U.quat_mul(env.agent.tcp.pose.q, U.quat_conj(q_before))
        d = torch.where(d[:, :1] < 0, -d, d)
        yaw = 2 * d[:, 3]
        if "target" in mode:
            assert (yaw - (-0.3)).abs().max() < 0.02, yaw
        else:
            assert (yaw < -0.05).all() and (yaw > -0.31).all(), yaw
    if "target" in mode:
        st = env.agent.controller.get_state()
        assert st["arm"]["target_pose"].shape == (n, 7)


def test_absolute_ee_pose_mode_reaches_the_commanded_pose():
    """pd_ee_pose (panda.py:125-136, pd_ee_pose.py:252-263): the action is the target pose itself -- position and XYZ Euler angles in the
    robot's root frame, not normalised; held for a second the TCP arrives there, orientation included."""
    import maniskill_b200 as ms
    from emu_world import EmuBackendWorld
    from maniskill_b200 import utils as U
    env = ms.make("PickCube-v1", num_envs=2, obs_mode="state", control_mode="pd_ee_pose", world_factory=EmuBackendWorld)
    env.reset(seed=0)
    assert env.action_dim == 7 and np.allclose(env.single_action_space_low[:3], -2.0) and np.allclose(env.single_action_space_high[3:6], 2 * np.pi)
    arm = env.agent.controller.controllers["arm"]
    cur = arm.ee_pose_at_base
    target_p = cur.p + torch.tensor([[0.05, -0.04, -0.03], [-0.03, 0.05, 0.02]])
    yaw = torch.tensor([0.3, -0.2])
    dq = U.matrix_to_quat(U.euler_xyz_to_matrix(torch.stack([torch.zeros(2), torch.zeros(2), yaw], 1)))
    target_q = U.quat_mul(dq, cur.q)                                          # the current orientation turned about the root z axis
    eul = U.matrix_to_euler_xyz(U.quat_to_matrix(target_q))
    a = torch.hstack([target_p, eul, torch.ones(2, 1)])
    for _ in range(25):
        env.step(a)
    now = arm.ee_pose_at_base
    assert (now.p - target_p).abs().max() < 5e-3
    err = U.quat_mul(now.q, U.quat_conj(target_q))
    assert (2 * torch.atan2(err[:, 1:].norm(dim=1), err[:, 0].abs())).max() < 0.02
