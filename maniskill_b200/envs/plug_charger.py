"""PlugCharger-v1 -- mirror of mani_skill/envs/tasks/tabletop/plug_charger.py:20-330 on the b200sim backend.

Table scene + a two-pin charger (base box + two 1.5 mm thick pins) and a kinematic receptacle (five boxes around two slots with 0.5 mm
clearance per side) floating 10 cm above the table: plug the charger in.  Sparse / no reward only, like the reference.  State observation
9 + 9 + 7 + 3 x 7 = 46.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import building as actors
from .. import utils as U
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class PlugChargerEnv(PandaTabletopEnv):
    max_episode_steps = 200  # @register_env("PlugCharger-v1", max_episode_steps=200)
    SUPPORTED_ROBOTS = ("panda_wristcam",)
    default_robot_uids = "panda_wristcam"
    SUPPORTED_REWARD_MODES = ("none", "sparse")
    _base_size = [2e-2, 1.5e-2, 1.2e-2]
    _peg_size = [8e-3, 0.75e-3, 3.2e-3]
    _peg_gap = 7e-3
    _clearance = 5e-4
    _receptacle_size = [1e-2, 5e-2, 5e-2]

    # ---- plug_charger.py:82-110
    def _build_charger(self, peg_size, base_size, gap):
        builder = actors.scene_desc_builder(self.scene_desc)
        mat = actors.RenderMaterial()
        mat.set_base_color([1, 1, 1, 1])
        builder.add_box_collision(actors.Pose([peg_size[0], gap, 0]), peg_size)
        builder.add_box_visual(actors.Pose([peg_size[0], gap, 0]), peg_size, material=mat)
        builder.add_box_collision(actors.Pose([peg_size[0], -gap, 0]), peg_size)
        builder.add_box_visual(actors.Pose([peg_size[0], -gap, 0]), peg_size, material=mat)
        builder.add_box_collision(actors.Pose([-base_size[0], 0, 0]), base_size)
        builder.add_box_visual(actors.Pose([-base_size[0], 0, 0]), base_size, material=mat)
        builder.initial_pose = actors.Pose(p=[0, 0, self._base_size[2]])
        return builder.build(name="charger")

    # ---- plug_charger.py:112-165
    def _build_receptacle(self, peg_size, receptacle_size, gap):
        builder = actors.scene_desc_builder(self.scene_desc)
        sy = 0.5 * (receptacle_size[1] - peg_size[1] - gap)
        sz = 0.5 * (receptacle_size[2] - peg_size[2])
        dx, dy, dz = -receptacle_size[0], peg_size[1] + gap + sy, peg_size[2] + sz
        mat = actors.RenderMaterial()
        mat.set_base_color([1, 1, 1, 1])
        poses = [actors.Pose([dx, 0, dz]), actors.Pose([dx, 0, -dz]), actors.Pose([dx, dy, 0]), actors.Pose([dx, -dy, 0])]
        half_sizes = [[receptacle_size[0], receptacle_size[1], sz], [receptacle_size[0], receptacle_size[1], sz],
                      [receptacle_size[0], sy, receptacle_size[2]], [receptacle_size[0], sy, receptacle_size[2]]]
        for pose, half_size in zip(poses, half_sizes):
            builder.add_box_collision(pose, half_size)
            builder.add_box_visual(pose, half_size, material=mat)
        pose, half_size = actors.Pose([-receptacle_size[0], 0, 0]), [receptacle_size[0], gap - peg_size[1], peg_size[2]]
        builder.add_box_collision(pose, half_size)
        builder.add_box_visual(pose, half_size, material=mat)
        gold = actors.RenderMaterial()
        gold.set_base_color([0xDB / 255, 0xB5 / 255, 0x39 / 255, 1.0])     # sapien_utils.hex2rgba("#DBB539")
        half_size = [receptacle_size[0], peg_size[1], peg_size[2]]
        builder.add_box_visual(actors.Pose([-receptacle_size[0], -(gap * 0.5 + peg_size[1]), 0]), half_size, material=gold)
        builder.add_box_visual(actors.Pose([-receptacle_size[0], gap * 0.5 + peg_size[1], 0]), half_size, material=gold)
        builder.initial_pose = actors.Pose(p=[0, 0, 0.1])
        return builder.build_kinematic(name="receptacle")

    # ---- plug_charger.py:170-190
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        self._build_charger(self._peg_size, self._base_size, self._peg_gap)
        self._build_receptacle([self._peg_size[0], self._peg_size[1] + self._clearance, self._peg_size[2] + self._clearance], self._receptacle_size, self._peg_gap)

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.charger = self.scene.actors["charger"]
        self.receptacle = self.scene.actors["receptacle"]
        self.goal_pose = None

    # ---- plug_charger.py:60-80 (the human render camera rides on the receptacle)
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.6], [-0.1, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0,
                     mount=None)] + self._robot_sensor_configs()

    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.3, 0.4, 0.1], [0, 0, 0]), width=512, height=512, fov=1, near=0.01, far=100.0,
                     mount=("actor", "receptacle"))]

    # ---- table/scene_builder.py:104-127 + plug_charger.py:192-251
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        # the task then overrides the wrist-camera rest pose: last arm joint at +pi/4, noise from the torch stream (plug_charger.py:197-222)
        qpos = torch.tensor([0.0, np.pi / 8, 0, -np.pi * 5 / 8, 0, np.pi * 3 / 4, np.pi / 4, 0.04, 0.04], device=dev)
        qpos = torch.normal(0, self.robot_init_qpos_noise, (b, len(qpos)), device=dev) + qpos
        qpos[:, -2:] = 0.04
        self.agent.robot.set_qpos(qpos)
        self.agent.robot.set_pose(Pose.create(np.array([-0.615, 0, 0, 1, 0, 0, 0], dtype=np.float32), dev))
        lo, hi = torch.tensor([-0.1, -0.2], device=dev), torch.tensor([-0.01 - self._peg_size[0] * 2, 0.2], device=dev)
        pos = torch.zeros((b, 3), device=dev)
        pos[:, :2] = torch.rand((b, 2), device=dev) * (hi - lo) + lo
        pos[:, 2] = self._base_size[2]
        ori = U.random_quaternions(b, device=dev, lock_x=True, lock_y=True, bounds=(-np.pi / 3, np.pi / 3))
        self.charger.set_pose(Pose.create_from_pq(pos, ori))
        lo, hi = torch.tensor([0.01, -0.1], device=dev), torch.tensor([0.1, 0.1], device=dev)
        pos = torch.zeros((b, 3), device=dev)
        pos[:, :2] = torch.rand((b, 2), device=dev) * (hi - lo) + lo
        pos[:, 2] = 0.1
        ori = U.random_quaternions(b, device=dev, lock_x=True, lock_y=True, bounds=(np.pi - np.pi / 8, np.pi + np.pi / 8))
        self.receptacle.set_pose(Pose.create_from_pq(pos, ori))
        self.goal_pose = self.receptacle.pose * Pose.create(np.array([0, 0, 0, *U.euler2quat(0, 0, np.pi)], dtype=np.float32), dev)

    # ---- plug_charger.py:253-255
    @property
    def charger_base_pose(self):
        return self.charger.pose * Pose.create(np.array([-self._base_size[0], 0, 0, 1, 0, 0, 0], dtype=np.float32), self.device)

    # ---- plug_charger.py:257-275
    def _compute_distance(self):
        obj_pose = self.charger.pose
        obj_to_goal_dist = torch.linalg.norm(self.goal_pose.p - obj_pose.p, axis=1)
        rel = U.quat_mul(U.quat_conj(self.goal_pose.q), obj_pose.q)
        rel = torch.where(rel[..., :1] < 0, -rel, rel)        # quaternion_multiply standardises the real part
        angle = torch.linalg.norm(U.quat_to_axis_angle(rel), axis=1)
        return obj_to_goal_dist, torch.min(angle, np.pi * 2 - angle)

    # ---- plug_charger.py:277-285
    def evaluate(self):
        dist, angle = self._compute_distance()
        return dict(obj_to_goal_dist=dist, obj_to_goal_angle=angle, success=(dist <= 5e-3) & (angle <= 0.2))

    # ---- plug_charger.py:287-296
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if self.obs_mode_struct.use_state:
            obs.update(charger_pose=self.charger.pose.raw_pose, receptacle_pose=self.receptacle.pose.raw_pose, goal_pose=self.goal_pose.raw_pose)
        return obs
