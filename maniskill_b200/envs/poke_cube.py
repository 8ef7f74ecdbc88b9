"""PokeCube-v1 -- mirror of mani_skill/envs/tasks/tabletop/poke_cube.py:25-283 on the b200sim backend.

Table scene + a 4 cm cube, a 24 x 5 x 5 cm peg lying along +x in front of it and a non-colliding goal disc beyond the cube: hold the peg
and poke the cube into the goal.  State observation 9 + 9 + 7 (tcp) + 7 + 7 + 3 * 5 = 54.  Task logic on the torch path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import building as actors
from .. import utils as U
from ..model import SHAPE_BOX, ActorRec, ShapeRec, pose7
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class PokeCubeEnv(PandaTabletopEnv):
    max_episode_steps = 50  # @register_env("PokeCube-v1", max_episode_steps=50)
    cube_half_size = 0.02
    peg_half_width = 0.025
    peg_half_length = 0.12
    goal_radius = 0.05

    # ---- poke_cube.py:81-124
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        h = self.cube_half_size
        actors.build_cube(self.scene_desc, half_size=h, color=[1, 0, 0, 1], name="cube", body_type="dynamic", initial_pose=actors.Pose(p=[1, 0, h]))
        blue = np.array([12, 42, 160, 255]) / 255
        actors.build_twocolor_peg(self.scene_desc, length=self.peg_half_length, width=self.peg_half_width, color_1=blue, color_2=blue, name="peg",
                                  body_type="dynamic", initial_pose=actors.Pose(p=[0, 0, self.peg_half_width]))
        # red/white target of the reference (thin visual cylinders, kinematic): a flat square of the same extent stands in for it
        self.scene_desc.add_actor(ActorRec("goal_region", "kinematic",
                                           [ShapeRec(SHAPE_BOX, pose7(), np.array([1e-5, self.goal_radius, self.goal_radius]), color=(0.9, 0.1, 0.1, 1), collide=False)],
                                           pose7()))

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.cube = self.scene.actors["cube"]
        self.peg = self.scene.actors["peg"]
        self.goal_region = self.scene.actors["goal_region"]
        self.peg_head_offsets = Pose.create_from_pq(torch.tensor([[self.peg_half_length, 0.0, 0.0]], device=self.device), device=self.device)

    # ---- poke_cube.py:126-132 (the position is offset in the WORLD frame, as in the reference)
    @property
    def peg_head_pos(self):
        return self.peg.pose.p + self.peg_head_offsets.p

    @property
    def peg_head_pose(self):
        return self.peg.pose * self.peg_head_offsets

    # ---- poke_cube.py:67-70
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.6], [-0.1, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=None)] + self._robot_sensor_configs()

    # ---- poke_cube.py:72-75
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, 0.7, 0.6], [0.2, 0.2, 0.35]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:68-103 + poke_cube.py:134-172
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        peg_xyz = torch.rand((b, 3), device=dev) * 0.2 - 0.1
        peg_xyz[:, 2] = self.peg_half_width
        self.peg.set_pose(Pose.create_from_pq(peg_xyz, device=dev))
        cube_xyz = torch.rand((b, 3), device=dev) * 0.2 - 0.1
        cube_xyz[:, 0] = peg_xyz[:, 0] + self.peg_half_length + 0.1
        cube_xyz[:, 2] = self.cube_half_size
        cube_q = U.random_quaternions(b, device=dev, lock_x=True, lock_y=True, lock_z=False, bounds=(-np.pi / 6, np.pi / 6))
        self.cube.set_pose(Pose.create_from_pq(cube_xyz, cube_q))
        goal_xyz = cube_xyz + torch.tensor([0.05 + self.goal_radius, 0, 0], device=dev)
        goal_xyz[:, 2] = 1e-3
        q = torch.tensor(U.euler2quat(0, np.pi / 2, 0), dtype=torch.float32, device=dev)
        self.goal_region.set_pose(Pose.create_from_pq(goal_xyz, q[None].expand(b, 4), device=dev))

    # ---- poke_cube.py:174-190
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            peg_p, cube_p = self.peg.pose.p, self.cube.pose.p
            obs.update(cube_pose=self.cube.pose.raw_pose, peg_pose=self.peg.pose.raw_pose, goal_pos=peg_p,
                       tcp_to_peg_pos=peg_p - self.agent.tcp.pose.p, peg_to_cube_pos=cube_p - peg_p,
                       cube_to_goal_pos=self.goal_region.pose.p - cube_p, peghead_to_cube_pos=self.peg_head_pos - cube_p)
        return obs

    # ---- poke_cube.py:192-233
    def evaluate(self):
        cube_p = self.cube.pose.p
        is_cube_placed = torch.linalg.norm(cube_p[..., :2] - self.goal_region.pose.p[..., :2], axis=1) < self.goal_radius
        peg_euler = U.matrix_to_euler_xyz(U.quat_to_matrix(self.peg_head_pose.q))
        cube_euler = U.matrix_to_euler_xyz(U.quat_to_matrix(self.cube.pose.q))
        angle_diff = torch.abs(peg_euler[:, 2] - cube_euler[:, 2])
        head_to_cube_dist = torch.linalg.norm(self.peg_head_pos[..., :2] - cube_p[..., :2], axis=1)
        is_peg_cube_fit = (angle_diff < 0.05) & (head_to_cube_dist <= self.cube_half_size + 0.005)
        return {"success": is_cube_placed & self.agent.is_static(0.2), "is_cube_placed": is_cube_placed, "is_peg_cube_fit": is_peg_cube_fit,
                "is_peg_grasped": self.agent.is_grasping(self.peg), "angle_diff": angle_diff, "head_to_cube_dist": head_to_cube_dist}

    # ---- poke_cube.py:235-276
    def compute_dense_reward(self, obs, action, info):
        tcp_to_peg_dist = torch.linalg.norm(self.agent.tcp.pose.p - self.peg.pose.p, axis=1)
        reward = 2 * (1 - torch.tanh(5.0 * tcp_to_peg_dist))
        align_reward = 1 - torch.tanh(5.0 * info["angle_diff"])
        close_reward = 1 - torch.tanh(5.0 * info["head_to_cube_dist"])
        is_peg_grasped = info["is_peg_grasped"] & (tcp_to_peg_dist < 0.01)
        reward = torch.where(is_peg_grasped, 4 + close_reward + align_reward, reward)
        cube_to_goal_dist = torch.linalg.norm(self.goal_region.pose.p - self.cube.pose.p, axis=1)
        reward = torch.where(info["is_peg_cube_fit"] & is_peg_grasped, 7 + (1 - torch.tanh(5 * cube_to_goal_dist)), reward)
        static_reward = 1 - torch.tanh(5 * torch.linalg.norm(self.agent.robot.get_qvel()[..., :-2], axis=1))
        reward = reward + static_reward * info["is_cube_placed"]
        return torch.where(info["success"], 10.0, reward)

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 10.0
