"""Robot agents + joint-space controllers -- host-side mirror of mani_skill/agents/base_agent.py:46,
mani_skill/agents/controllers/pd_joint_pos.py (PDJointPosController / PDJointPosMimicController) and
mani_skill/agents/robots/panda/panda.py (Panda).  Same action layout, bounds, scaling and mimic rule as the
reference; the output of ``set_action`` is written straight into the backend's ``target_qpos`` buffer
(articulation.py:873-896) and picked up by the fused substep kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import utils as U
from .structs import Articulation, Pose


class PDJointPosController:
    """pd_joint_pos.py:15-101.  ``lower/upper`` None => joint limits, ``use_delta`` => target = qpos + action."""
    sets_target_qpos = True
    sets_target_qvel = False

    def __init__(self, articulation: Articulation, joint_names: List[str], lower, upper, use_delta=False, use_target=False,
                 normalize_action=True):
        self.articulation = articulation
        self.scene = articulation.scene
        self.device = self.scene.device
        self.joint_names = joint_names
        self.active_joint_indices = torch.tensor([articulation.dof_names.index(n) for n in joint_names], dtype=torch.int64, device=self.device)
        self.use_delta, self.use_target, self.normalize_action = use_delta, use_target, normalize_action
        lim = articulation.qlimits[0, self.active_joint_indices].cpu().numpy().copy()
        if lower is not None:
            lim[:, 0] = lower
        if upper is not None:
            lim[:, 1] = upper
        self._set_bounds(lim)
        self._target_qpos = None
        self._start_qpos = None

    def _set_bounds(self, lim):
        self.action_low = torch.tensor(lim[:, 0], dtype=torch.float32, device=self.device)
        self.action_high = torch.tensor(lim[:, 1], dtype=torch.float32, device=self.device)
        self.action_dim = lim.shape[0]

    @property
    def qpos(self):
        return self.articulation.qpos[..., self.active_joint_indices]

    def reset(self, env_idx=None):
        q = self.qpos.clone()
        if self._target_qpos is None or env_idx is None:
            self._start_qpos, self._target_qpos = q.clone(), q.clone()
        else:
            self._start_qpos[env_idx] = q[env_idx]
            self._target_qpos[env_idx] = q[env_idx]

    def _preprocess_action(self, action):
        if self.normalize_action:
            action = U.clip_and_scale_action(action, self.action_low, self.action_high)
        return action

    def set_drive_targets(self, targets):
        self.articulation.set_joint_drive_targets(targets, self.active_joint_indices)

    def set_action(self, action):
        action = self._preprocess_action(action)
        self._start_qpos = self.qpos
        if self.use_delta:
            self._target_qpos = (self._target_qpos if self.use_target else self._start_qpos) + action
        else:
            self._target_qpos = torch.broadcast_to(action, self._start_qpos.shape).clone()
        self.set_drive_targets(self._target_qpos)

    def get_state(self):
        return {"target_qpos": self._target_qpos} if self.use_target else {}


class PDJointPosMimicController(PDJointPosController):
    """pd_joint_pos.py:129-237: one action per control joint, mimic joints copy multiplier*q + offset."""

    def __init__(self, articulation, joint_names, lower, upper, mimic: Dict[str, dict], **kw):
        super().__init__(articulation, joint_names, lower, upper, **kw)
        mimic_idx, ctrl_idx = [], []
        for mimic_name, data in mimic.items():
            mimic_idx.append(joint_names.index(mimic_name))
            ctrl_idx.append(joint_names.index(data["joint"]))
        self.mimic_joint_indices = torch.tensor(mimic_idx, dtype=torch.int64, device=self.device)
        self.mimic_control_joint_indices = torch.tensor(ctrl_idx, dtype=torch.int64, device=self.device)
        self.control_joint_indices = torch.unique(self.mimic_control_joint_indices)
        self._multiplier = torch.tensor([d.get("multiplier", 1.0) for d in mimic.values()], dtype=torch.float32, device=self.device)
        self._offset = torch.tensor([d.get("offset", 0.0) for d in mimic.values()], dtype=torch.float32, device=self.device)
        lim = np.stack([self.action_low.cpu().numpy(), self.action_high.cpu().numpy()], 1)[self.control_joint_indices.cpu().numpy()]
        self._set_bounds(lim)

    def set_action(self, action):
        action = self._preprocess_action(action)
        self._start_qpos = self.qpos
        self._target_qpos = self._target_qpos.clone()
        if self.use_delta:
            base = self._target_qpos if self.use_target else self._start_qpos
            self._target_qpos[:, self.control_joint_indices] = base[:, self.control_joint_indices] + action
        else:
            self._target_qpos[:, self.control_joint_indices] = action
        self._target_qpos[:, self.mimic_joint_indices] = (
            self._target_qpos[:, self.mimic_control_joint_indices] * self._multiplier[None, :] + self._offset[None, :])
        self.set_drive_targets(self._target_qpos)


class PDJointVelController:
    """pd_joint_vel.py:13-48: the (clipped, scaled) action is the velocity target of the joints' drives; the drives are built with
    stiffness 0 (the gains are part of the compiled model: the task passes the control mode to the scene description)."""
    sets_target_qpos = False
    sets_target_qvel = True
    use_target = False

    def __init__(self, articulation: Articulation, joint_names: List[str], lower, upper, normalize_action=True):
        self.articulation = articulation
        self.scene = articulation.scene
        self.device = self.scene.device
        self.joint_names = joint_names
        self.active_joint_indices = torch.tensor([articulation.dof_names.index(n) for n in joint_names], dtype=torch.int64, device=self.device)
        n = len(joint_names)
        self.normalize_action = normalize_action
        self.action_low = torch.tensor(np.broadcast_to(np.float32(lower), n).copy(), device=self.device)
        self.action_high = torch.tensor(np.broadcast_to(np.float32(upper), n).copy(), device=self.device)
        self.action_dim = n

    @property
    def qpos(self):
        return self.articulation.qpos[..., self.active_joint_indices]

    def reset(self, env_idx=None):
        """A reset sub-scene starts with zero velocity targets."""
        rows = self.articulation._rows if env_idx is None else self.articulation._rows[env_idx]
        self.scene.world.target_qvel[rows[:, None], self.active_joint_indices[None, :]] = 0.0
        self.scene._dirty |= self.scene.BUF_TARGET_QVEL

    def _preprocess_action(self, action):
        return U.clip_and_scale_action(action, self.action_low, self.action_high) if self.normalize_action else action

    def set_action(self, action):
        self.articulation.set_joint_drive_velocity_targets(self._preprocess_action(action), self.active_joint_indices)

    def get_state(self):
        return {}


class PDJointPosVelController(PDJointPosController):
    """pd_joint_pos_vel.py:11-69: the action is [position part | velocity part]; the position half goes through the PDJointPos rules
    (absolute / delta / target-delta), the velocity half becomes the drives' velocity targets."""
    sets_target_qvel = True

    def __init__(self, articulation: Articulation, joint_names: List[str], lower, upper, vel_lower=-1.0, vel_upper=1.0, **kw):
        super().__init__(articulation, joint_names, lower, upper, **kw)
        n = len(joint_names)
        lim = np.stack([np.concatenate([self.action_low.cpu().numpy(), np.broadcast_to(np.float32(vel_lower), n)]),
                        np.concatenate([self.action_high.cpu().numpy(), np.broadcast_to(np.float32(vel_upper), n)])], 1)
        self._set_bounds(lim)
        self._target_qvel = None

    def reset(self, env_idx=None):
        super().reset(env_idx)
        if self._target_qvel is None or env_idx is None:
            self._target_qvel = torch.zeros_like(self.qpos)
        else:
            self._target_qvel[env_idx] = 0.0
        rows = self.articulation._rows if env_idx is None else self.articulation._rows[env_idx]
        self.scene.world.target_qvel[rows[:, None], self.active_joint_indices[None, :]] = 0.0
        self.scene._dirty |= self.scene.BUF_TARGET_QVEL

    def set_action(self, action):
        action = self._preprocess_action(action)
        nq = action.shape[1] // 2
        self._start_qpos = self.qpos
        if self.use_delta:
            self._target_qpos = (self._target_qpos if self.use_target else self._start_qpos) + action[:, :nq]
        else:
            self._target_qpos = torch.broadcast_to(action[:, :nq], self._start_qpos.shape).clone()
        self.set_drive_targets(self._target_qpos)
        self._target_qvel = action[:, nq:]
        self.articulation.set_joint_drive_velocity_targets(self._target_qvel, self.active_joint_indices)


class PDEEPosController(PDJointPosController):
    """pd_ee_pose.py:25-146, GPU-simulation semantics: frame "root_translation", delta actions.  Without a virtual target the
    (clipped and scaled) action IS the end-effector displacement in the root frame and goes straight into one damped least-squares
    IK step; with `use_target` the displacement is applied to the previous target pose and the IK step closes the gap between
    that target and the current end-effector pose (pd_ee_pose.py:101-133, utils/kinematics.py:197-260)."""
    pos_only = True

    def __init__(self, articulation: Articulation, joint_names: List[str], pos_lower, pos_upper, robot: dict, ee_link: str,
                 rot_lower=None, rot_upper=None, use_delta=True, use_target=False, normalize_action=True,
                 solver_config: Optional[dict] = None):
        from .kinematics import Kinematics
        super().__init__(articulation, joint_names, None, None, use_delta=use_delta, use_target=use_target, normalize_action=normalize_action)
        self.kinematics = Kinematics(robot, ee_link, articulation.dof_names, joint_names, self.device, articulation=articulation)
        self.ee_link = articulation.links_map[ee_link]
        self.root_link = articulation.root
        self.solver_config = dict(type="levenberg_marquardt", alpha=1.0) if solver_config is None else solver_config
        low = np.broadcast_to(np.float32(pos_lower), 3)
        high = np.broadcast_to(np.float32(pos_upper), 3)
        self.rot_lower = rot_lower
        if not self.pos_only:
            low = np.hstack([low, np.broadcast_to(np.float32(rot_lower), 3)])
            high = np.hstack([high, np.broadcast_to(np.float32(rot_upper), 3)])
        self._set_bounds(np.stack([low, high], 1))
        self._target_pose = None

    @property
    def ee_pose_at_base(self) -> Pose:
        return self.root_link.pose.inv() * self.ee_link.pose

    def reset(self, env_idx=None):
        super().reset(env_idx)
        if self.use_target:
            cur = self.ee_pose_at_base.raw_pose
            if self._target_pose is None or env_idx is None:
                self._target_pose = Pose(cur.clone())
            else:
                self._target_pose.raw_pose[env_idx] = cur[env_idx]

    def compute_target_pose(self, prev: Pose, action) -> Pose:
        """pd_ee_pose.py:85-99 with frame root_translation: keep the rotation, translate in the root frame (`delta_pose * prev`: the
        pose product standardises the quaternion to a non-negative real part, pose.py:199)."""
        return Pose(torch.hstack([prev.p + action[:, :3], torch.where(prev.q[..., :1] < 0, -prev.q, prev.q)]))

    def _delta_from_target(self, target: Pose, current: Pose):
        """utils/kinematics.py:218-241: translation difference and XYZ Euler angles of target.q * current.q^-1."""
        dq = U.quat_mul(target.q, U.quat_conj(current.q))
        dq = torch.where(dq[..., :1] < 0, -dq, dq)
        return torch.hstack([target.p - current.p, U.matrix_to_euler_xyz(U.quat_to_matrix(dq))])

    def set_action(self, action):
        action = self._preprocess_action(action)
        self._start_qpos = self.qpos
        if self.use_target:
            self._target_pose = self.compute_target_pose(self._target_pose, action)
            delta = self._delta_from_target(self._target_pose, self.ee_pose_at_base)
        elif not self.use_delta:
            # absolute target in the root frame (pd_ee_pose.py:95-99,252-263 -> utils/kinematics.py:211-231: the 6-vector becomes a pose,
            # the IK step closes the gap to the current end-effector pose)
            delta = self._delta_from_target(self.compute_target_pose(None, action), self.ee_pose_at_base)
        else:
            delta = torch.hstack([action, torch.zeros((action.shape[0], 3), device=self.device)]) if self.pos_only else action
        self._target_qpos = self.kinematics.compute_ik(delta, self.articulation.get_qpos(), self.solver_config)
        self.set_drive_targets(self._target_qpos)

    def get_state(self):
        return {"target_pose": self._target_pose.raw_pose} if self.use_target else {}


class PDEEPoseController(PDEEPosController):
    """pd_ee_pose.py:203-263, frame "root_translation:root_aligned_body_rotation": the last three action entries are a rotation
    vector-like increment (clipped by norm, scaled by `rot_lower` exactly as pd_ee_pose.py:229-239 does), applied as XYZ Euler
    angles in the root frame."""
    pos_only = False

    def _preprocess_action(self, action):
        if not self.normalize_action:
            return action
        pos = U.clip_and_scale_action(action[:, :3], self.action_low[:3], self.action_high[:3])
        rot = action[:, 3:]
        norm = torch.linalg.norm(rot, dim=1, keepdim=True)
        rot = torch.where(norm > 1, rot / norm.clamp(min=1e-12), rot) * float(np.broadcast_to(self.rot_lower, 3)[0])
        return torch.hstack([pos, rot])

    def compute_target_pose(self, prev: Pose, action) -> Pose:
        dq = U.matrix_to_quat(U.euler_xyz_to_matrix(action[:, 3:6]))
        if not self.use_delta:   # pd_ee_pose.py:252-263: position and XYZ Euler angles of the target, root frame
            return Pose(torch.hstack([action[:, :3], dq]))
        q = U.quat_mul(dq, prev.q)
        q = torch.where(q[..., :1] < 0, -q, q)
        return Pose(torch.hstack([prev.p + action[:, :3], q]))


class CombinedController:
    """base_controller.py:305-347 `DictController` with balanced action concatenation."""

    def __init__(self, controllers: Dict[str, PDJointPosController]):
        self.controllers = controllers
        self.action_mapping = {}
        d = 0
        for k, c in controllers.items():
            self.action_mapping[k] = (d, d + c.action_dim)
            d += c.action_dim
        self.action_dim = d

    @property
    def sets_target_qpos(self) -> bool:
        return any(c.sets_target_qpos for c in self.controllers.values())

    @property
    def sets_target_qvel(self) -> bool:
        return any(c.sets_target_qvel for c in self.controllers.values())

    def before_simulation_step(self):
        """Interpolated targets are not part of this build: nothing changes between the simulation steps of a control step."""

    def reset(self, env_idx=None):
        for c in self.controllers.values():
            c.reset(env_idx)

    def set_action(self, action):
        for k, c in self.controllers.items():
            a, b = self.action_mapping[k]
            c.set_action(action[:, a:b])

    def get_state(self):
        out = {}
        for k, c in self.controllers.items():
            s = c.get_state()
            if s:
                out[k] = s
        return out


class Panda:
    """mani_skill/agents/robots/panda/panda.py:16-269 (uid 'panda'; PandaWristCam = panda_v3 urdf)."""
    uid = "panda"
    arm_joint_names = [f"panda_joint{i}" for i in range(1, 8)]
    gripper_joint_names = ["panda_finger_joint1", "panda_finger_joint2"]
    ee_link_name = "panda_hand_tcp"
    SUPPORTED_CONTROL_MODES = ("pd_joint_delta_pos", "pd_joint_pos", "pd_ee_delta_pos", "pd_ee_delta_pose", "pd_ee_pose", "pd_joint_target_delta_pos",
                               "pd_ee_target_delta_pos", "pd_ee_target_delta_pose", "pd_joint_vel", "pd_joint_pos_vel", "pd_joint_delta_pos_vel")
    arm_stiffness, arm_damping, arm_force_limit = 1e3, 1e2, 100.0          # panda.py:68-74

    @classmethod
    def drive_gains(cls, control_mode: Optional[str]):
        """(stiffness, damping, force limit) of the arm drives under a control mode (panda.py:76-170: the velocity controller builds its
        drives without stiffness, pd_joint_vel.py:29-36); the gains are part of the compiled scene."""
        return (0.0, cls.arm_damping, cls.arm_force_limit) if control_mode == "pd_joint_vel" else (cls.arm_stiffness, cls.arm_damping, cls.arm_force_limit)
    robot_asset = "panda_v2"

    def __init__(self, scene, name="panda"):
        self.scene = scene
        self.device = scene.device
        self.uid = name                       # "panda" | "panda_wristcam" (panda_wristcam.py:12-16 only changes uid, urdf and the camera)
        self.robot: Articulation = scene.articulations[name]
        lm = self.robot.links_map
        self.finger1_link = lm["panda_leftfinger"]
        self.finger2_link = lm["panda_rightfinger"]
        self.tcp = lm[self.ee_link_name]
        self.controller = None
        self.control_mode = None

    def set_control_mode(self, control_mode: Optional[str] = None):
        if control_mode is None:
            control_mode = self.SUPPORTED_CONTROL_MODES[0]  # first key of the controller dict (panda.py:187-190)
        if control_mode not in self.SUPPORTED_CONTROL_MODES:
            raise NotImplementedError(f"control mode {control_mode} is not available in this build "
                                      f"(supported: {self.SUPPORTED_CONTROL_MODES})")
        self.control_mode = control_mode
        gripper = PDJointPosMimicController(self.robot, self.gripper_joint_names, -0.01, 0.04,
                                            mimic={"panda_finger_joint2": {"joint": "panda_finger_joint1"}})
        if control_mode.startswith("pd_ee_"):  # panda.py:103-141: pos +-0.1, rot +-0.1, ee link panda_hand_tcp
            from .model import load_robot
            cls = PDEEPoseController if control_mode.endswith("pose") else PDEEPosController
            if control_mode == "pd_ee_pose":      # panda.py:125-136: absolute pose target, +-2 m, rotation +-2 pi, not normalised
                arm = cls(self.robot, self.arm_joint_names, -2.0, 2.0, load_robot(self.robot_asset), self.ee_link_name, rot_lower=-2 * np.pi,
                          rot_upper=2 * np.pi, use_delta=False, normalize_action=False)
            else:
                arm = cls(self.robot, self.arm_joint_names, -0.1, 0.1, load_robot(self.robot_asset), self.ee_link_name, rot_lower=-0.1, rot_upper=0.1,
                          use_target="target" in control_mode)
        elif control_mode == "pd_joint_delta_pos":
            arm = PDJointPosController(self.robot, self.arm_joint_names, -0.1, 0.1, use_delta=True)
        elif control_mode == "pd_joint_target_delta_pos":
            arm = PDJointPosController(self.robot, self.arm_joint_names, -0.1, 0.1, use_delta=True, use_target=True)
        elif control_mode == "pd_joint_vel":                   # panda.py:144-150
            arm = PDJointVelController(self.robot, self.arm_joint_names, -1.0, 1.0)
        elif control_mode == "pd_joint_pos_vel":               # panda.py:153-161
            arm = PDJointPosVelController(self.robot, self.arm_joint_names, None, None, normalize_action=False)
        elif control_mode == "pd_joint_delta_pos_vel":         # panda.py:162-170
            arm = PDJointPosVelController(self.robot, self.arm_joint_names, -0.1, 0.1, use_delta=True)
        else:
            arm = PDJointPosController(self.robot, self.arm_joint_names, None, None, normalize_action=False)
        self.controller = CombinedController(dict(arm=arm, gripper=gripper))

    def action_bounds(self):
        lows, highs = [], []
        for c in self.controller.controllers.values():
            if c.normalize_action:
                lows.append(-np.ones(c.action_dim, dtype=np.float32))
                highs.append(np.ones(c.action_dim, dtype=np.float32))
            else:
                lows.append(c.action_low.cpu().numpy())
                highs.append(c.action_high.cpu().numpy())
        return np.concatenate(lows), np.concatenate(highs)

    def reset(self, init_qpos=None):
        """base_agent.py:300-320: zero velocity / force, optional qpos."""
        if init_qpos is not None:
            self.robot.set_qpos(init_qpos)
        self.robot.set_qvel(torch.zeros(self.robot.dof, device=self.device))
        self.robot.set_qf(torch.zeros(self.robot.dof, device=self.device))

    def controller_reset(self, env_idx=None):
        self.controller.reset(env_idx)
        # the drive keeps holding the reset configuration until the first action (controller.reset -> targets = qpos)
        full = self.robot.qpos.clone()
        rows = self.robot._rows if env_idx is None else self.robot._rows[env_idx]
        self.scene.world.target_qpos[rows, :self.robot.dof] = full if env_idx is None else full[env_idx]
        self.scene._dirty |= self.scene.BUF_TARGET_QPOS
        self.scene._gpu_apply_all()

    def set_action(self, action):
        self.controller.set_action(action)

    def before_simulation_step(self):
        """base_agent.py:332-334."""
        self.controller.before_simulation_step()

    def get_proprioception(self):
        obs = dict(qpos=self.robot.get_qpos(), qvel=self.robot.get_qvel())
        cs = self.controller.get_state()
        if cs:
            obs["controller"] = cs
        return obs

    def is_grasping(self, obj, min_force=0.5, max_angle=85):
        """panda.py:237-265."""
        l_f = self.scene.get_pairwise_contact_forces(self.finger1_link, obj)
        r_f = self.scene.get_pairwise_contact_forces(self.finger2_link, obj)
        lforce = torch.linalg.norm(l_f, axis=1)
        rforce = torch.linalg.norm(r_f, axis=1)
        ldirection = self.finger1_link.pose.to_transformation_matrix()[..., :3, 1]
        rdirection = -self.finger2_link.pose.to_transformation_matrix()[..., :3, 1]
        langle = U.compute_angle_between(ldirection, l_f)
        rangle = U.compute_angle_between(rdirection, r_f)
        lflag = torch.logical_and(lforce >= min_force, torch.rad2deg(langle) <= max_angle)
        rflag = torch.logical_and(rforce >= min_force, torch.rad2deg(rangle) <= max_angle)
        return torch.logical_and(lflag, rflag)

    def is_static(self, threshold: float = 0.2):
        qvel = self.robot.get_qvel()[..., :-2]
        return torch.max(torch.abs(qvel), 1)[0] <= threshold

    @property
    def tcp_pose(self) -> Pose:
        return self.tcp.pose

    @property
    def tcp_pos(self):
        return self.tcp.pose.p


class PDBaseForwardVelController:
    """mani_skill/agents/controllers/pd_base_vel.py:39-73: action = (forward velocity, yaw rate) in the robot frame ->
    velocity drive targets of the three virtual base joints (x, y, yaw)."""
    sets_target_qpos = False
    sets_target_qvel = True
    normalize_action = True
    use_target = False

    def __init__(self, articulation: Articulation, joint_names: List[str], lower, upper):
        self.articulation = articulation
        self.scene = articulation.scene
        self.device = self.scene.device
        self.active_joint_indices = torch.tensor([articulation.dof_names.index(n) for n in joint_names], dtype=torch.int64, device=self.device)
        self.action_low = torch.tensor(np.broadcast_to(lower, 2).astype(np.float32), device=self.device)
        self.action_high = torch.tensor(np.broadcast_to(upper, 2).astype(np.float32), device=self.device)
        self.action_dim = 2

    @property
    def qpos(self):
        return self.articulation.qpos[..., self.active_joint_indices]

    def reset(self, env_idx=None):
        pass

    def set_action(self, action):
        action = U.clip_and_scale_action(action, self.action_low, self.action_high)
        ori = self.qpos[:, 2]
        c, s = torch.cos(ori), torch.sin(ori)
        vel = torch.stack([c * action[:, 0], s * action[:, 0]], 1)  # rot(ori) @ (forward, 0)
        self.articulation.set_joint_drive_velocity_targets(torch.hstack([vel, action[:, 1:]]), self.active_joint_indices)

    def get_state(self):
        return {}


class Fetch:
    """mani_skill/agents/robots/fetch/fetch.py:26-410 (uid 'fetch'): 7-dof arm, mimic gripper, head/torso, velocity-driven base."""
    uid = "fetch"
    arm_joint_names = ["shoulder_pan_joint", "shoulder_lift_joint", "upperarm_roll_joint", "elbow_flex_joint", "forearm_roll_joint",
                       "wrist_flex_joint", "wrist_roll_joint"]
    gripper_joint_names = ["l_gripper_finger_joint", "r_gripper_finger_joint"]
    body_joint_names = ["head_pan_joint", "head_tilt_joint", "torso_lift_joint"]
    base_joint_names = ["root_x_axis_joint", "root_y_axis_joint", "root_z_rotation_joint"]
    ee_link_name = "gripper_link"
    SUPPORTED_CONTROL_MODES = ("pd_joint_delta_pos", "pd_joint_pos")

    def __init__(self, scene, name="fetch"):
        self.scene = scene
        self.device = scene.device
        self.robot: Articulation = scene.articulations[name]
        lm = self.robot.links_map
        self.finger1_link, self.finger2_link = lm["l_gripper_finger_link"], lm["r_gripper_finger_link"]
        self.tcp = lm[self.ee_link_name]
        self.base_link = lm["base_link"]
        self.controller = None
        self.control_mode = None

    def set_control_mode(self, control_mode: Optional[str] = None):
        if control_mode is None:
            control_mode = self.SUPPORTED_CONTROL_MODES[0]
        if control_mode not in self.SUPPORTED_CONTROL_MODES:
            raise NotImplementedError(f"control mode {control_mode} is not available in this build (supported: {self.SUPPORTED_CONTROL_MODES})")
        self.control_mode = control_mode
        if control_mode == "pd_joint_delta_pos":
            arm = PDJointPosController(self.robot, self.arm_joint_names, -0.1, 0.1, use_delta=True)
        else:
            arm = PDJointPosController(self.robot, self.arm_joint_names, None, None, normalize_action=False)
        gripper = PDJointPosMimicController(self.robot, self.gripper_joint_names, -0.01, 0.05,
                                            mimic={"r_gripper_finger_joint": {"joint": "l_gripper_finger_joint"}})
        body = PDJointPosController(self.robot, self.body_joint_names, -0.1, 0.1, use_delta=True)
        base = PDBaseForwardVelController(self.robot, self.base_joint_names, [-1, -3.14], [1, 3.14])
        self.controller = CombinedController(dict(arm=arm, gripper=gripper, body=body, base=base))

    action_bounds = Panda.action_bounds
    reset = Panda.reset
    set_action = Panda.set_action
    before_simulation_step = Panda.before_simulation_step
    get_proprioception = Panda.get_proprioception

    def controller_reset(self, env_idx=None):
        Panda.controller_reset(self, env_idx)
        rows = self.robot._rows if env_idx is None else self.robot._rows[env_idx]
        self.scene.world.target_qvel[rows, :self.robot.dof] = 0.0
        self.scene._dirty |= self.scene.BUF_TARGET_QVEL
        self.scene._gpu_apply_all()

    def is_grasping(self, obj, min_force=0.5, max_angle=85):
        """fetch.py:338-368 (finger opening directions are mirrored with respect to the Panda)."""
        l_f = self.scene.get_pairwise_contact_forces(self.finger1_link, obj)
        r_f = self.scene.get_pairwise_contact_forces(self.finger2_link, obj)
        ldirection = -self.finger1_link.pose.to_transformation_matrix()[..., :3, 1]
        rdirection = self.finger2_link.pose.to_transformation_matrix()[..., :3, 1]
        lflag = torch.logical_and(torch.linalg.norm(l_f, axis=1) >= min_force, torch.rad2deg(U.compute_angle_between(ldirection, l_f)) <= max_angle)
        rflag = torch.logical_and(torch.linalg.norm(r_f, axis=1) >= min_force, torch.rad2deg(U.compute_angle_between(rdirection, r_f)) <= max_angle)
        return torch.logical_and(lflag, rflag)

    def is_static(self, threshold: float = 0.2, base_threshold: float = 0.05):
        """fetch.py:370-375 (the reference orders qvel base-first; here the joints are selected by name)."""
        names = self.robot.dof_names
        base = torch.tensor([names.index(n) for n in self.base_joint_names], device=self.device)
        fingers = set(self.gripper_joint_names) | set(self.base_joint_names)
        body = torch.tensor([i for i, n in enumerate(names) if n not in fingers], device=self.device)
        qv = self.robot.get_qvel()
        return torch.all(qv[:, body] <= threshold, dim=1) & torch.all(qv[:, base] <= base_threshold, dim=1)

    @property
    def tcp_pose(self) -> Pose:
        return self.tcp.pose
