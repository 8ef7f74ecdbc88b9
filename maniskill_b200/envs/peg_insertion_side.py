"""PegInsertionSide-v1 -- mirror of mani_skill/envs/tasks/tabletop/peg_insertion_side.py:50-360 on the b200sim backend.

Heterogeneous sub-scenes: every env has its own peg (dynamic box, half size (L, r, r)) and its own box-with-hole
(kinematic actor of four boxes), drawn from `_batched_episode_rng` exactly like the reference (:114-120).  The reference
builds one actor per sub-scene and merges the views (`Actor.merge`, :185-191); here the prototype carries per-env
size/pose/mass override tables (include/b200sim.h `ov_*`), so the merged view is simply the body row.
Robot: `panda_wristcam` (panda_v3.urdf), cameras: `base_camera` + `hand_camera` mounted on `camera_link`.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import utils as U
from ..agents import Panda
from ..model import SHAPE_BOX, ActorRec, ShapeRec, pose7
from ..scenes import SQRT_HALF, TABLE_HEIGHT, add_table_scene, panda_articulation
from ..structs import Pose
from .base_env import BaseEnv


def hex2rgba(h):
    h = h.lstrip("#")
    c = [int(h[i:i + 2], 16) / 255.0 for i in (0, 2, 4)]
    # sapien_utils.hex2rgba converts sRGB -> linear by default (gamma 2.2)
    return tuple(float(x) ** 2.2 for x in c) + (1.0,)


class PegInsertionSideEnv(BaseEnv):
    graph_epilogue = True   # evaluate / observation / reward: pure tensor code, captured into a CUDA graph (base_env._GraphedEpilogue)
    max_episode_steps = 100
    _clearance = 0.003

    def __init__(self, *args, robot_uids="panda_wristcam", **kwargs):
        if robot_uids != "panda_wristcam":
            raise NotImplementedError("PegInsertionSide-v1 supports robot_uids='panda_wristcam' only (SUPPORTED_ROBOTS)")
        self.robot_uids = robot_uids
        kwargs.setdefault("fused", False)
        super().__init__(*args, **kwargs)

    def _load_agent_desc(self):
        art = panda_articulation("panda_wristcam", "panda_v3", (-0.615, 0, 0))
        for j in Panda.arm_joint_names:      # the drive gains follow the control mode the env is made with
            art.drive[j] = Panda.drive_gains(self._control_mode_arg)
        self.scene_desc.add_articulation(art)

    # ---- peg_insertion_side.py:109-191
    def _load_scene_desc(self):
        N = self.num_envs
        add_table_scene(self.scene_desc)
        rng = self._batched_episode_rng
        lengths = rng.uniform(0.085, 0.125)
        radii = rng.uniform(0.015, 0.025)
        centers = 0.5 * (lengths - radii)[:, None] * rng.uniform(-1, 1, size=(2,))
        self._lengths, self._radii, self._centers = lengths, radii, centers
        peg_sizes = np.stack([lengths, radii, radii], 1)
        half = np.stack([lengths / 2, radii, radii], 1)
        head_pose = np.tile(pose7(), (N, 1)); head_pose[:, 0] = lengths / 2
        tail_pose = np.tile(pose7(), (N, 1)); tail_pose[:, 0] = -lengths / 2
        peg = ActorRec("peg", "dynamic", [
            ShapeRec(SHAPE_BOX, pose7(), peg_sizes[0].copy(), per_env_size=peg_sizes, visual=False),
            ShapeRec(SHAPE_BOX, head_pose[0].copy(), half[0].copy(), per_env_size=half, per_env_pose=head_pose, collide=False, color=hex2rgba("#EC7357")),
            ShapeRec(SHAPE_BOX, tail_pose[0].copy(), half[0].copy(), per_env_size=half, per_env_pose=tail_pose, collide=False, color=hex2rgba("#EDF6F9")),
        ], pose7([0, 0, 0.1]))
        self.scene_desc.add_actor(peg)
        # _build_box_with_hole(inner_radius=r+clearance, outer_radius=L, depth=L, center)  (:19-47)
        inner, outer, depth = radii + self._clearance, lengths, lengths
        thick = (outer - inner) * 0.5
        hc = centers * 0.5
        sizes = [np.stack([depth, thick - hc[:, 0], outer], 1), np.stack([depth, thick + hc[:, 0], outer], 1),
                 np.stack([depth, outer, thick - hc[:, 1]], 1), np.stack([depth, outer, thick + hc[:, 1]], 1)]
        off = thick + inner
        z = np.zeros(N)
        poss = [np.stack([z, off + hc[:, 0], z], 1), np.stack([z, -off + hc[:, 0], z], 1), np.stack([z, z, off + hc[:, 1]], 1), np.stack([z, z, -off + hc[:, 1]], 1)]
        shapes = []
        for sz, ps in zip(sizes, poss):
            p7 = np.tile(pose7(), (N, 1)); p7[:, :3] = ps
            shapes.append(ShapeRec(SHAPE_BOX, p7[0].copy(), sz[0].copy(), per_env_size=sz, per_env_pose=p7, color=hex2rgba("#FFD289")))
        self.scene_desc.add_actor(ActorRec("box_with_hole", "kinematic", shapes, pose7([0, 1, 0.1])))

    def _after_build(self):
        dev = self.device
        self.agent = Panda(self.scene, "panda_wristcam")
        self.table = self.scene.actors["table-workspace"]
        self.peg = self.scene.actors["peg"]
        self.box = self.scene.actors["box_with_hole"]
        self.peg_half_sizes = torch.tensor(np.stack([self._lengths, self._radii, self._radii], 1), dtype=torch.float32, device=dev)
        off = torch.zeros((self.num_envs, 3), device=dev)
        off[:, 0] = self.peg_half_sizes[:, 0]
        self.peg_head_offsets = Pose.create_from_pq(p=off, device=dev)
        hole = torch.zeros((self.num_envs, 3), device=dev)
        hole[:, 1:] = torch.tensor(self._centers, dtype=torch.float32, device=dev)
        self.box_hole_offsets = Pose.create_from_pq(p=hole, device=dev)
        self.box_hole_radii = torch.tensor(self._radii + self._clearance, dtype=torch.float32, device=dev)

    # ---- :96-99 and agents/robots/panda/panda_wristcam.py:19-32
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0, -0.3, 0.2], [0, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=None),
                dict(uid="hand_camera", pose=pose7(), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=("panda_wristcam", "camera_link"))]

    # ---- :101-104
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.5, -0.5, 0.8], [0.05, -0.1, 0.4]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- :193-249
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self.table.set_pose(Pose.create(pose7([-0.12, 0, -TABLE_HEIGHT], [SQRT_HALF, 0, 0, SQRT_HALF]), dev))
        # table_scene.initialize for panda_wristcam draws one (b, 9) normal and resets the robot (table/scene_builder.py:104-127)
        qpos = np.array([0.0, np.pi / 8, 0, -np.pi * 5 / 8, 0, np.pi * 3 / 4, -np.pi / 4, 0.04, 0.04])
        q = self._episode_rng.normal(0, 0.02, (b, 9)) + qpos
        q[:, -2:] = 0.04
        self.agent.reset(torch.tensor(q, dtype=torch.float32, device=dev))
        self.agent.robot.set_pose(Pose.create(pose7([-0.615, 0, 0]), dev))
        xy = torch.rand((b, 2), device=dev) * (torch.tensor([0.1, 0.0], device=dev) - torch.tensor([-0.1, -0.3], device=dev)) + torch.tensor([-0.1, -0.3], device=dev)
        pos = torch.zeros((b, 3), device=dev)
        pos[:, :2] = xy
        pos[:, 2] = self.peg_half_sizes[env_idx, 2]
        quat = U.random_quaternions(b, dev, lock_x=True, lock_y=True, bounds=(np.pi / 2 - np.pi / 3, np.pi / 2 + np.pi / 3))
        self.peg.set_pose(Pose.create_from_pq(pos, quat))
        xy = torch.rand((b, 2), device=dev) * (torch.tensor([0.05, 0.4], device=dev) - torch.tensor([-0.05, 0.2], device=dev)) + torch.tensor([-0.05, 0.2], device=dev)
        pos = torch.zeros((b, 3), device=dev)
        pos[:, :2] = xy
        pos[:, 2] = self.peg_half_sizes[env_idx, 0]
        quat = U.random_quaternions(b, dev, lock_x=True, lock_y=True, bounds=(np.pi / 2 - np.pi / 8, np.pi / 2 + np.pi / 8))
        self.box.set_pose(Pose.create_from_pq(pos, quat))
        # the task then redraws the robot configuration (:232-249)
        q = self._episode_rng.normal(0, 0.02, (b, 9)) + qpos
        q[:, -2:] = 0.04
        self.agent.robot.set_qpos(torch.tensor(q, dtype=torch.float32, device=dev))
        self.agent.robot.set_pose(Pose.create(pose7([-0.615, 0, 0]), dev))

    # ---- :251-287.  The three derived poses are used by evaluate, the observation and the reward of the same step: computed once per
    # simulation state (the reference recomputes them on every access and notes they could be cached, :262-266).
    def _memo(self, name, fn):
        version = getattr(self, "_state_version", None)
        if version is None:
            return fn()
        cache = self.__dict__.setdefault("_pose_cache", {})
        hit = cache.get(name)
        if hit is None or hit[0] != version:
            hit = (version, fn())
            cache[name] = hit
        return hit[1]

    @property
    def peg_head_pose(self):
        return PegInsertionSideEnv._memo(self, "peg_head", lambda: self.peg.pose * self.peg_head_offsets)

    @property
    def box_hole_pose(self):
        return PegInsertionSideEnv._memo(self, "box_hole", lambda: self.box.pose * self.box_hole_offsets)

    @property
    def goal_pose(self):
        return PegInsertionSideEnv._memo(self, "goal", lambda: self.box_hole_pose * self.peg_head_offsets.inv())

    def has_peg_inserted(self):
        p = (self.box_hole_pose.inv() * self.peg_head_pose).p
        x_flag = -0.015 <= p[:, 0]
        y_flag = (-self.box_hole_radii <= p[:, 1]) & (p[:, 1] <= self.box_hole_radii)
        z_flag = (-self.box_hole_radii <= p[:, 2]) & (p[:, 2] <= self.box_hole_radii)
        return x_flag & y_flag & z_flag, p

    def evaluate(self):
        success, p = self.has_peg_inserted()
        return dict(success=success, peg_head_pos_at_hole=p)

    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            obs.update(peg_pose=self.peg.pose.raw_pose, peg_half_size=self.peg_half_sizes, box_hole_pose=self.box_hole_pose.raw_pose,
                       box_hole_radius=self.box_hole_radii)
        return obs

    # ---- :289-360
    def compute_dense_reward(self, obs, action, info):
        gripper_pos = self.agent.tcp.pose.p
        if getattr(self, "_grasp_offset", None) is None:
            self._grasp_offset = Pose.create(pose7([-0.06, 0, 0]), self.device)
        tgt = self.peg.pose * self._grasp_offset
        gripper_to_peg_dist = torch.linalg.norm(gripper_pos - tgt.p, axis=1)
        reaching_reward = 1 - torch.tanh(4.0 * gripper_to_peg_dist)
        is_grasped = self.agent.is_grasping(self.peg, max_angle=20)
        reward = reaching_reward + is_grasped
        goal_inv = self.goal_pose.inv()
        peg_head_wrt_goal = goal_inv * self.peg_head_pose
        d_head = torch.linalg.norm(peg_head_wrt_goal.p[:, 1:], axis=1)
        peg_wrt_goal = goal_inv * self.peg.pose
        d_peg = torch.linalg.norm(peg_wrt_goal.p[:, 1:], axis=1)
        pre_insertion_reward = 3 * (1 - torch.tanh(0.5 * (d_head + d_peg) + 4.5 * torch.maximum(d_head, d_peg)))
        reward = reward + pre_insertion_reward * is_grasped
        pre_inserted = (d_head < 0.01) & (d_peg < 0.01)
        inside = self.box_hole_pose.inv() * self.peg_head_pose
        insertion_reward = 5 * (1 - torch.tanh(5.0 * torch.linalg.norm(inside.p, axis=1)))
        reward = reward + insertion_reward * (is_grasped & pre_inserted)
        reward = torch.where(info["success"], 10.0, reward)  # masked assignment without the nonzero() sync
        return reward

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs, action, info) / 10
