"""PullCube-v1 -- mirror of mani_skill/envs/tasks/tabletop/pull_cube.py:20-152 on the b200sim backend: the PushCube-v1 scene with the
goal disc behind the cube (towards the robot), no height condition in the success test, observation 9 + 9 + 7 + 3 + 7 = 35."""
from __future__ import annotations

import numpy as np
import torch

from .. import utils as U
from ..model import pose7
from ..structs import Pose
from .push_cube import PushCubeEnv


class PullCubeEnv(PushCubeEnv):
    max_episode_steps = 50  # @register_env("PullCube-v1", max_episode_steps=50)

    # ---- pull_cube.py:83-103
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, :2] = torch.rand((b, 2), device=dev) * 0.2 - 0.1
        xyz[:, 2] = self.cube_half_size
        self.obj.set_pose(Pose.create_from_pq(xyz, device=dev))
        target = xyz - torch.tensor([0.1 + self.goal_radius, 0, 0], device=dev)
        target[:, 2] = 1e-3
        q = torch.tensor(U.euler2quat(0, np.pi / 2, 0), dtype=torch.float32, device=dev)
        self.goal_region.set_pose(Pose.create_from_pq(target, q[None].expand(b, 4), device=dev))

    # ---- pull_cube.py:105-115
    def evaluate(self):
        return {"success": torch.linalg.norm(self.obj.pose.p[..., :2] - self.goal_region.pose.p[..., :2], axis=1) < self.goal_radius}

    # ---- pull_cube.py:117-126
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose, goal_pos=self.goal_region.pose.p)
        if "state" in self.obs_mode:
            obs.update(obj_pose=self.obj.pose.raw_pose)
        return obs

    # ---- pull_cube.py:128-152: close the gripper behind the cube (its far side) and drag it
    def compute_dense_reward(self, obs, action, info):
        pull_p = self.obj.pose.p + torch.tensor([self.cube_half_size + 2 * 0.005, 0, 0], device=self.obj.pose.p.device)
        dist = torch.linalg.norm(pull_p - self.agent.tcp.pose.p, axis=1)
        reward = 1 - torch.tanh(5 * dist)
        obj_to_goal = torch.linalg.norm(self.obj.pose.p[..., :2] - self.goal_region.pose.p[..., :2], axis=1)
        reward = reward + (1 - torch.tanh(5 * obj_to_goal)) * (dist < 0.01)
        return torch.where(info["success"], 3.0, reward)

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 3.0
