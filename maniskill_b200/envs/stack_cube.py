"""StackCube-v1 -- mirror of mani_skill/envs/tasks/tabletop/stack_cube.py:18-200 on the b200sim backend.

Robot `panda_wristcam` (the reference default), table scene, two 4 cm cubes placed by the reference's rejection sampler with the
same torch.rand call order; state observation 9 + 9 + 7 + 7 + 7 + 3 + 3 + 3 = 48; success = cube A on cube B, static and released.
Task logic on the torch path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import utils as U
from ..model import SHAPE_BOX, ActorRec, ShapeRec, pose7
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class StackCubeEnv(PandaTabletopEnv):
    max_episode_steps = 50  # @register_env("StackCube-v1", max_episode_steps=50)

    default_robot_uids = "panda_wristcam"   # stack_cube.py:36-41; "panda" is accepted too (SUPPORTED_ROBOTS)

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("fused", False)
        super().__init__(*args, **kwargs)

    # ---- stack_cube.py:57-77
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        h = np.array([0.02, 0.02, 0.02])
        self.scene_desc.add_actor(ActorRec("cubeA", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), h, color=(1, 0, 0, 1))], pose7([0, 0, 0.1])))
        self.scene_desc.add_actor(ActorRec("cubeB", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), h, color=(0, 1, 0, 1))], pose7([1, 0, 0.1])))

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.cubeA = self.scene.actors["cubeA"]
        self.cubeB = self.scene.actors["cubeB"]
        self.cube_half_size = torch.tensor([0.02] * 3, dtype=torch.float32, device=self.device)

    # ---- stack_cube.py:45-48 and agents/robots/panda/panda_wristcam.py:19-32
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.6], [-0.1, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=None),
                ] + self._robot_sensor_configs()

    # ---- stack_cube.py:49-52
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:104-127 + stack_cube.py:79-113
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, 2] = 0.02
        xy = torch.rand((b, 2), device=dev) * 0.2 - 0.1
        sampler = U.UniformPlacementSampler([[-0.1, -0.2], [0.1, 0.2]], b, device=dev)
        radius = torch.linalg.norm(torch.tensor([0.02, 0.02])) + 0.001
        cubeA_xy = xy + sampler.sample(radius, 100)
        cubeB_xy = xy + sampler.sample(radius, 100, verbose=False)
        xyz[:, :2] = cubeA_xy
        self.cubeA.set_pose(Pose.create_from_pq(xyz.clone(), U.random_quaternions(b, device=dev, lock_x=True, lock_y=True)))
        xyz[:, :2] = cubeB_xy
        self.cubeB.set_pose(Pose.create_from_pq(xyz, U.random_quaternions(b, device=dev, lock_x=True, lock_y=True)))

    # ---- stack_cube.py:115-135 (Actor.is_static: utils/structs/actor.py:220-227)
    def evaluate(self):
        offset = self.cubeA.pose.p - self.cubeB.pose.p
        xy_flag = torch.linalg.norm(offset[..., :2], axis=1) <= torch.linalg.norm(self.cube_half_size[:2]) + 0.005
        z_flag = torch.abs(offset[..., 2] - self.cube_half_size[..., 2] * 2) <= 0.005
        is_cubeA_on_cubeB = torch.logical_and(xy_flag, z_flag)
        is_cubeA_static = torch.logical_and(torch.linalg.norm(self.cubeA.linear_velocity, axis=1) <= 1e-2,
                                            torch.linalg.norm(self.cubeA.angular_velocity, axis=1) <= 0.5)
        is_cubeA_grasped = self.agent.is_grasping(self.cubeA)
        success = is_cubeA_on_cubeB & is_cubeA_static & (~is_cubeA_grasped)
        return {"is_cubeA_grasped": is_cubeA_grasped, "is_cubeA_on_cubeB": is_cubeA_on_cubeB, "is_cubeA_static": is_cubeA_static, "success": success}

    # ---- stack_cube.py:137-148
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            tcp_p = self.agent.tcp.pose.p
            obs.update(cubeA_pose=self.cubeA.pose.raw_pose, cubeB_pose=self.cubeB.pose.raw_pose, tcp_to_cubeA_pos=self.cubeA.pose.p - tcp_p,
                       tcp_to_cubeB_pos=self.cubeB.pose.p - tcp_p, cubeA_to_cubeB_pos=self.cubeB.pose.p - self.cubeA.pose.p)
        return obs

    # ---- stack_cube.py:150-195
    def compute_dense_reward(self, obs, action, info):
        tcp_p = self.agent.tcp.pose.p
        cubeA_pos, cubeB_pos = self.cubeA.pose.p, self.cubeB.pose.p
        reward = 2 * (1 - torch.tanh(5 * torch.linalg.norm(tcp_p - cubeA_pos, axis=1)))
        goal_xyz = torch.hstack([cubeB_pos[:, 0:2], (cubeB_pos[:, 2] + self.cube_half_size[2] * 2)[:, None]])
        place_reward = 1 - torch.tanh(5.0 * torch.linalg.norm(goal_xyz - cubeA_pos, axis=1))
        grasped = info["is_cubeA_grasped"]
        reward = torch.where(grasped, 4 + place_reward, reward)
        gripper_width = self.agent.robot.get_qlimits()[0, -1, 1] * 2  # hard-coded with panda, like the reference
        ungrasp_reward = torch.sum(self.agent.robot.get_qpos()[:, -2:], axis=1) / gripper_width
        ungrasp_reward = torch.where(grasped, ungrasp_reward, torch.ones_like(ungrasp_reward))
        v = torch.linalg.norm(self.cubeA.linear_velocity, axis=1)
        av = torch.linalg.norm(self.cubeA.angular_velocity, axis=1)
        static_reward = 1 - torch.tanh(v * 10 + av)
        reward = torch.where(info["is_cubeA_on_cubeB"], 6 + (ungrasp_reward + static_reward) / 2.0, reward)
        return torch.where(info["success"], 8.0, reward)

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 8
