"""PushCube-v1 -- mirror of mani_skill/envs/tasks/tabletop/push_cube.py:36-241 on the b200sim backend.

Same scene (table scene + 4 cm cube + non-colliding goal disc), randomisation (same torch.rand call order under the same seed),
state observation (9 + 9 + 7 + 3 + 7 = 35), success test and dense reward as the reference task; the task logic runs on the torch path
(only the pick family has a fused epilogue kernel).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import utils as U
from ..model import SHAPE_BOX, ActorRec, ShapeRec, pose7
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class PushCubeEnv(PandaTabletopEnv):
    max_episode_steps = 50  # @register_env("PushCube-v1", max_episode_steps=50)
    goal_radius = 0.1
    cube_half_size = 0.02

    # ---- push_cube.py:105-141
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        h = self.cube_half_size
        self.scene_desc.add_actor(ActorRec("cube", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([h, h, h]), color=(12 / 255, 42 / 255, 160 / 255, 1))],
                                           pose7([0, 0, h])))
        # red/white target of the reference (a thin cylinder, visual only, kinematic): a flat square of the same extent stands in for it
        self.scene_desc.add_actor(ActorRec("goal_region", "kinematic",
                                           [ShapeRec(SHAPE_BOX, pose7(), np.array([1e-5, self.goal_radius, self.goal_radius]), color=(0.9, 0.1, 0.1, 1), collide=False)],
                                           pose7([0, 0, 1e-3])))

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.obj = self.scene.actors["cube"]
        self.goal_region = self.scene.actors["goal_region"]

    # ---- push_cube.py:84-92
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.6], [-0.1, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=None)] + self._robot_sensor_configs()

    # ---- push_cube.py:93-99 (PullCube-v1: pull_cube.py:49-52, the same camera)
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:68-103 + push_cube.py:143-177
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, :2] = torch.rand((b, 2), device=dev) * 0.2 - 0.1
        xyz[:, 2] = self.cube_half_size
        self.obj.set_pose(Pose.create_from_pq(xyz, device=dev))
        target = xyz + torch.tensor([0.1 + self.goal_radius, 0, 0], device=dev)
        target[:, 2] = 1e-3
        q = torch.tensor(U.euler2quat(0, np.pi / 2, 0), dtype=torch.float32, device=dev)
        self.goal_region.set_pose(Pose.create_from_pq(target, q[None].expand(b, 4), device=dev))

    # ---- push_cube.py:179-192
    def evaluate(self):
        is_obj_placed = (torch.linalg.norm(self.obj.pose.p[..., :2] - self.goal_region.pose.p[..., :2], axis=1) < self.goal_radius) & (
            self.obj.pose.p[..., 2] < self.cube_half_size + 5e-3)
        return {"success": is_obj_placed}

    # ---- push_cube.py:194-207
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            obs.update(goal_pos=self.goal_region.pose.p, obj_pose=self.obj.pose.raw_pose)
        return obs

    # ---- push_cube.py:209-241
    def compute_dense_reward(self, obs, action, info):
        push_p = self.obj.pose.p + torch.tensor([-self.cube_half_size - 0.005, 0, 0], device=self.obj.pose.p.device)
        tcp_to_push_pose_dist = torch.linalg.norm(push_p - self.agent.tcp.pose.p, axis=1)
        reward = 1 - torch.tanh(5 * tcp_to_push_pose_dist)
        reached = tcp_to_push_pose_dist < 0.01
        obj_to_goal_dist = torch.linalg.norm(self.obj.pose.p[..., :2] - self.goal_region.pose.p[..., :2], axis=1)
        place_reward = 1 - torch.tanh(5 * obj_to_goal_dist)
        reward = reward + place_reward * reached
        z_deviation = torch.abs(self.obj.pose.p[..., 2] - self.cube_half_size)
        z_reward = 1 - torch.tanh(5 * z_deviation)
        reward = reward + place_reward * z_reward * reached
        return torch.where(info["success"], 4.0, reward)  # masked assignment without the nonzero() sync

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 4.0
