"""LiftPegUpright-v1 -- mirror of mani_skill/envs/tasks/tabletop/lift_peg_upright.py:20-137 on the b200sim backend.

Table scene + a 24 x 5 x 5 cm two-colour peg lying along the world x axis (rolled a quarter turn about it); success when the peg stands on one of its small faces.
State observation 9 + 9 + 7 (tcp) + 7 (peg) = 32.  Task logic on the torch path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import building as actors
from .. import utils as U
from ..model import SHAPE_BOX, ShapeRec, pose7
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


def twocolor_peg_shapes(half_length, half_width, color_1, color_2):
    """actors/common.py:230-261 build_twocolor_peg: one collision box, two visual halves along x."""
    half = np.array([half_length / 2, half_width, half_width])
    return [ShapeRec(SHAPE_BOX, pose7(), np.array([half_length, half_width, half_width]), visual=False),
            ShapeRec(SHAPE_BOX, pose7([-half_length / 2, 0, 0]), half, collide=False, color=tuple(color_1)),
            ShapeRec(SHAPE_BOX, pose7([half_length / 2, 0, 0]), half, collide=False, color=tuple(color_2))]


class LiftPegUprightEnv(PandaTabletopEnv):
    max_episode_steps = 50  # @register_env("LiftPegUpright-v1", max_episode_steps=50)
    peg_half_width = 0.025
    peg_half_length = 0.12

    # ---- lift_peg_upright.py:54-72
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        actors.build_twocolor_peg(self.scene_desc, length=self.peg_half_length, width=self.peg_half_width, color_1=np.array([176, 14, 14, 255]) / 255,
                                  color_2=np.array([12, 42, 160, 255]) / 255, name="peg", body_type="dynamic", initial_pose=actors.Pose(p=[0, 0, 0.1]))

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.peg = self.scene.actors["peg"]

    # ---- lift_peg_upright.py:44-47
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.6], [-0.1, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=None)] + self._robot_sensor_configs()

    # ---- lift_peg_upright.py:49-52
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:68-103 + lift_peg_upright.py:74-86
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, :2] = torch.rand((b, 2), device=dev) * 0.2 - 0.1
        xyz[:, 2] = self.peg_half_width
        q = torch.tensor(U.euler2quat(np.pi / 2, 0, 0), dtype=torch.float32, device=dev)
        self.peg.set_pose(Pose.create_from_pq(xyz, q[None].expand(b, 4), device=dev))

    # ---- lift_peg_upright.py:88-99 (the third XYZ Euler angle, as the reference code reads it)
    def evaluate(self):
        euler = U.matrix_to_euler_xyz(U.quat_to_matrix(self.peg.pose.q))
        is_peg_upright = torch.abs(torch.abs(euler[:, 2]) - np.pi / 2) < 0.08
        close_to_table = torch.abs(self.peg.pose.p[:, 2] - self.peg_half_length) < 0.005
        return {"success": is_peg_upright & close_to_table}

    # ---- lift_peg_upright.py:101-109
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            obs.update(obj_pose=self.peg.pose.raw_pose)
        return obs

    # ---- lift_peg_upright.py:111-137
    def compute_dense_reward(self, obs, action, info):
        rot_vec = U.quat_to_matrix(self.peg.pose.q)[..., :, 0]  # peg long axis in the world
        reward = rot_vec[:, 2].abs()
        z_dist = torch.abs(self.peg.pose.p[:, 2] - self.peg_half_length)
        reward = reward + 1 - torch.tanh(5 * z_dist)
        to_grip_dist = torch.linalg.norm(self.peg.pose.p - self.agent.tcp.pose.p, axis=1)
        reaching_rew = torch.where(self.agent.is_grasping(self.peg), 1.0, 1 - torch.tanh(5 * to_grip_dist)) / 5
        reward = reward + reaching_rew
        return torch.where(info["success"], 3.0, reward)

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 3.0
