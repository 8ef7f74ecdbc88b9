"""PullCubeTool-v1 -- mirror of mani_skill/envs/tasks/tabletop/pull_cube_tool.py:20-282 on the b200sim backend.

Table scene + an L-shaped tool (two boxes: a 20 cm handle of half density and a hook) within reach and a 4 cm cube out of reach: grasp the
tool and drag the cube towards the robot with it.  State observation 9 + 9 + 7 (tcp) + 7 (cube) + 7 (tool) = 39.  Task logic on the torch path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import building as actors
from .. import utils as U
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class PullCubeToolEnv(PandaTabletopEnv):
    max_episode_steps = 100  # @register_env("PullCubeTool-v1", max_episode_steps=100)
    SUPPORTED_ROBOTS = ("panda",)
    goal_radius = 0.3
    cube_half_size = 0.02
    handle_length = 0.2
    hook_length = 0.05
    width = 0.05
    height = 0.05
    cube_size = 0.02
    arm_reach = 0.35

    # ---- pull_cube_tool.py:94-130
    def _build_l_shaped_tool(self, handle_length, hook_length, width, height):
        builder = actors.scene_desc_builder(self.scene_desc)
        mat = actors.RenderMaterial()
        mat.set_base_color([1, 0, 0, 1])
        mat.metallic, mat.roughness, mat.specular = 1.0, 0.0, 1.0
        builder.add_box_collision(actors.Pose([handle_length / 2, 0, 0]), [handle_length / 2, width / 2, height / 2], density=500)
        builder.add_box_visual(actors.Pose([handle_length / 2, 0, 0]), [handle_length / 2, width / 2, height / 2], material=mat)
        builder.add_box_collision(actors.Pose([handle_length - hook_length / 2, width, 0]), [hook_length / 2, width, height / 2])
        builder.add_box_visual(actors.Pose([handle_length - hook_length / 2, width, 0]), [hook_length / 2, width, height / 2], material=mat)
        return builder.build(name="l_shape_tool")

    # ---- pull_cube_tool.py:132-153
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        actors.build_cube(self.scene_desc, half_size=self.cube_half_size, color=np.array([12, 42, 160, 255]) / 255, name="cube", body_type="dynamic")
        self._build_l_shaped_tool(self.handle_length, self.hook_length, self.width, self.height)

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.cube = self.scene.actors["cube"]
        self.l_shape_tool = self.scene.actors["l_shape_tool"]

    # ---- pull_cube_tool.py:61-92
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.5], [-0.1, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0,
                     mount=None)] + self._robot_sensor_configs()

    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:68-103 + pull_cube_tool.py:155-188
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        tool_xyz = torch.zeros((b, 3), device=dev)
        tool_xyz[:, :2] = -torch.rand((b, 2), device=dev) * 0.2 - 0.1
        tool_xyz[:, 2] = self.height / 2
        self.l_shape_tool.set_pose(Pose.create_from_pq(tool_xyz, device=dev))
        cube_xyz = torch.zeros((b, 3), device=dev)
        cube_xyz[:, 0] = self.arm_reach + torch.rand(b, device=dev) * self.handle_length - 0.3
        cube_xyz[:, 1] = torch.rand(b, device=dev) * 0.3 - 0.25
        cube_xyz[:, 2] = self.cube_size / 2 + 0.015
        cube_q = U.random_quaternions(b, device=dev, lock_x=True, lock_y=True, lock_z=False, bounds=(-np.pi / 6, np.pi / 6))
        self.cube.set_pose(Pose.create_from_pq(cube_xyz, cube_q))

    # ---- pull_cube_tool.py:190-201
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if self.obs_mode_struct.use_state:
            obs.update(cube_pose=self.cube.pose.raw_pose, tool_pose=self.l_shape_tool.pose.raw_pose)
        return obs

    # ---- pull_cube_tool.py:203-228 (the scalar progress entries and the embedded reward are part of the reference's info dict)
    def evaluate(self):
        cube_pos = self.cube.pose.p
        robot_base_pos = self.agent.robot.get_links()[0].pose.p
        cube_pulled_close = torch.linalg.norm(cube_pos[:, :2] - robot_base_pos[:, :2], dim=1) < 0.6
        workspace_center = robot_base_pos.clone()
        workspace_center[:, 0] += self.arm_reach * 0.1
        cube_to_workspace_dist = torch.linalg.norm(cube_pos - workspace_center, dim=1)
        progress = 1 - torch.tanh(3.0 * cube_to_workspace_dist)
        return {"success": cube_pulled_close, "success_once": cube_pulled_close, "success_at_end": cube_pulled_close, "cube_progress": progress.mean(),
                "cube_distance": cube_to_workspace_dist.mean(), "reward": self.compute_normalized_dense_reward(None, None, {"success": cube_pulled_close})}

    # ---- pull_cube_tool.py:230-272
    def compute_dense_reward(self, obs, action, info):
        dev = self.device
        tcp_pos, cube_pos, tool_pos = self.agent.tcp.pose.p, self.cube.pose.p, self.l_shape_tool.pose.p
        robot_base_pos = self.agent.robot.get_links()[0].pose.p
        tcp_to_tool_dist = torch.linalg.norm(tcp_pos - (tool_pos + torch.tensor([0.02, 0, 0], device=dev)), dim=1)
        reaching_reward = 2.0 * (1 - torch.tanh(5.0 * tcp_to_tool_dist))
        is_grasping = self.agent.is_grasping(self.l_shape_tool, max_angle=20)
        ideal_hook_pos = cube_pos + torch.tensor([-(self.hook_length + self.cube_half_size), -0.067, 0], device=dev)
        tool_positioning_dist = torch.linalg.norm(tool_pos - ideal_hook_pos, dim=1)
        positioning_reward = 1.5 * (1 - torch.tanh(3.0 * tool_positioning_dist))
        workspace_target = robot_base_pos + torch.tensor([0.05, 0, 0], device=dev)
        cube_to_workspace_dist = torch.linalg.norm(cube_pos - workspace_target, dim=1)
        initial_dist = torch.linalg.norm(torch.tensor([self.arm_reach + 0.1, 0, self.cube_size / 2], device=dev) - workspace_target, dim=1)
        pulling_reward = 3.0 * ((initial_dist - cube_to_workspace_dist) / initial_dist) * (tool_positioning_dist < 0.05)
        reward = reaching_reward + 2.0 * is_grasping + positioning_reward * is_grasping + pulling_reward * is_grasping
        reward = reward - 2.0 * (cube_pos[:, 0] > (self.arm_reach + 0.15))
        if "success" in info:
            reward = reward + 5.0 * info["success"]
        return reward

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 5.0
