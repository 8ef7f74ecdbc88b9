"""PlaceSphere-v1 -- mirror of mani_skill/envs/tasks/tabletop/place_sphere.py:24-307 on the b200sim backend.

Table scene + a 2 cm sphere and a shallow kinematic bin (a bottom plate and four rims, five boxes): pick the sphere up and set it down
in the bin.  State observation 9 + 9 + 1 + 7 (tcp) + 3 (bin) + 7 (sphere) + 3 = 39.  Task logic on the torch path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import building as actors
from .. import utils as U
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class PlaceSphereEnv(PandaTabletopEnv):
    max_episode_steps = 50  # @register_env("PlaceSphere-v1", max_episode_steps=50)
    radius = 0.02
    inner_side_half_len = 0.02
    short_side_half_size = 0.0025
    block_half_size = [short_side_half_size, 2 * short_side_half_size + inner_side_half_len, 2 * short_side_half_size + inner_side_half_len]
    edge_block_half_size = [short_side_half_size, 2 * short_side_half_size + inner_side_half_len, 2 * short_side_half_size]

    # ---- place_sphere.py:96-130
    def _build_bin(self):
        builder = actors.scene_desc_builder(self.scene_desc)
        dx = dy = self.block_half_size[1] - self.block_half_size[0]
        dz = self.edge_block_half_size[2] + self.block_half_size[0]
        e = self.edge_block_half_size
        poses = [actors.Pose([0, 0, 0]), actors.Pose([-dx, 0, dz]), actors.Pose([dx, 0, dz]), actors.Pose([0, -dy, dz]), actors.Pose([0, dy, dz])]
        half_sizes = [[self.block_half_size[1], self.block_half_size[2], self.block_half_size[0]], e, e, [e[1], e[0], e[2]], [e[1], e[0], e[2]]]
        for pose, half_size in zip(poses, half_sizes):
            builder.add_box_collision(pose, half_size)
            builder.add_box_visual(pose, half_size)
        return builder.build_kinematic(name="bin")

    # ---- place_sphere.py:135-148
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        actors.build_sphere(self.scene_desc, radius=self.radius, color=np.array([12, 42, 160, 255]) / 255, name="sphere", body_type="dynamic")
        self._build_bin()

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.obj = self.scene.actors["sphere"]
        self.bin = self.scene.actors["bin"]

    # ---- place_sphere.py:72-94
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.2], [-0.1, 0, 0]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0,
                     mount=None)] + self._robot_sensor_configs()

    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, -0.2, 0.2], [0.0, 0.0, 0.2]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:68-103 + place_sphere.py:150-184
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, 0] = (torch.rand((b, 1), device=dev) * 0.05 - 0.1)[:, 0]      # the quarter of the spawn area nearest to the robot
        xyz[:, 1] = (torch.rand((b, 1), device=dev) * 0.2 - 0.1)[:, 0]
        xyz[:, 2] = self.radius
        self.obj.set_pose(Pose.create_from_pq(xyz, device=dev))
        pos = torch.zeros((b, 3), device=dev)
        pos[:, 0] = torch.rand((b, 1), device=dev)[:, 0] * 0.1                   # the far half
        pos[:, 1] = torch.rand((b, 1), device=dev)[:, 0] * 0.2 - 0.1
        pos[:, 2] = self.block_half_size[0]
        self.bin.set_pose(Pose.create_from_pq(pos, device=dev))

    # ---- place_sphere.py:186-205
    def evaluate(self):
        offset = self.obj.pose.p - self.bin.pose.p
        xy_flag = torch.linalg.norm(offset[..., :2], axis=1) <= 0.005
        z_flag = torch.abs(offset[..., 2] - self.radius - self.block_half_size[0]) <= 0.005
        is_obj_on_bin = xy_flag & z_flag
        is_obj_static = self.obj.is_static(lin_thresh=1e-2, ang_thresh=0.5)
        is_obj_grasped = self.agent.is_grasping(self.obj)
        return {"is_obj_grasped": is_obj_grasped, "is_obj_on_bin": is_obj_on_bin, "is_obj_static": is_obj_static,
                "success": is_obj_on_bin & is_obj_static & (~is_obj_grasped)}

    # ---- place_sphere.py:207-218
    def _get_obs_extra(self, info: dict):
        obs = dict(is_grasped=info["is_obj_grasped"], tcp_pose=self.agent.tcp.pose.raw_pose, bin_pos=self.bin.pose.p)
        if "state" in self.obs_mode:
            obs.update(obj_pose=self.obj.pose.raw_pose, tcp_to_obj_pos=self.obj.pose.p - self.agent.tcp.pose.p)
        return obs

    # ---- place_sphere.py:220-265
    def compute_dense_reward(self, obs, action, info):
        obj_pos = self.obj.pose.p
        reward = 2 * (1 - torch.tanh(5 * torch.linalg.norm(self.agent.tcp.pose.p - obj_pos, axis=1)))
        bin_top_pos = self.bin.pose.p.clone()
        bin_top_pos[:, 2] = bin_top_pos[:, 2] + self.block_half_size[0] + self.radius
        place_reward = 1 - torch.tanh(5.0 * torch.linalg.norm(bin_top_pos - obj_pos, axis=1))
        is_obj_grasped = info["is_obj_grasped"]
        reward = torch.where(is_obj_grasped, 4 + place_reward, reward)
        gripper_width = self.agent.robot.get_qlimits()[0, -1, 1] * 2
        ungrasp_reward = torch.sum(self.agent.robot.get_qpos()[:, -2:], axis=1) / gripper_width
        ungrasp_reward = torch.where(is_obj_grasped, ungrasp_reward, 16.0)   # larger than the static terms, so that the gripper may close
        v = torch.linalg.norm(self.obj.linear_velocity, axis=1)
        av = torch.linalg.norm(self.obj.angular_velocity, axis=1)
        static_reward = 1 - torch.tanh(v * 10 + av)
        robot_static_reward = self.agent.is_static(0.2)
        reward = torch.where(info["is_obj_on_bin"], 6 + (ungrasp_reward + static_reward + robot_static_reward) / 3.0, reward)
        return torch.where(info["success"], 13.0, reward)

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 13.0
