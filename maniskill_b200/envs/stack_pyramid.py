"""StackPyramid-v1 -- mirror of mani_skill/envs/tasks/tabletop/stack_pyramid.py:23-214 on the b200sim backend.

Table scene + three 4 cm cubes placed without overlap (the reference's `UniformPlacementSampler` stream): put the red cube next to the green
one and the blue one on top of both.  Sparse / no reward only, like the reference.  State observation 9 + 9 + 7 + 3 x 7 + 6 x 3 = 64.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import building as actors
from .. import utils as U
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class StackPyramidEnv(PandaTabletopEnv):
    max_episode_steps = 250  # @register_env("StackPyramid-v1", max_episode_steps=250)
    default_robot_uids = "panda_wristcam"
    SUPPORTED_REWARD_MODES = ("none", "sparse")

    # ---- stack_pyramid.py:64-93
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        actors.build_cube(self.scene_desc, half_size=0.02, color=[1, 0, 0, 1], name="cubeA", initial_pose=actors.Pose(p=[0, 0, 0.2]))
        actors.build_cube(self.scene_desc, half_size=0.02, color=[0, 1, 0, 1], name="cubeB", initial_pose=actors.Pose(p=[1, 0, 0.2]))
        actors.build_cube(self.scene_desc, half_size=0.02, color=[0, 0, 1, 1], name="cubeC", initial_pose=actors.Pose(p=[-1, 0, 0.2]))

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.cubeA, self.cubeB, self.cubeC = (self.scene.actors[n] for n in ("cubeA", "cubeB", "cubeC"))
        self.cube_half_size = torch.tensor([0.02] * 3, dtype=torch.float32, device=self.device)

    # ---- stack_pyramid.py:54-62
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([0.3, 0, 0.4], [-0.05, 0, 0.1]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0,
                     mount=None)] + self._robot_sensor_configs()

    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:104-127 + stack_pyramid.py:95-145
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, 2] = 0.02
        xy = xyz[:, :2]
        sampler = U.UniformPlacementSampler([[-0.1, -0.2], [0.1, 0.2]], b, device=dev)
        radius = torch.linalg.norm(torch.tensor([0.02, 0.02]))
        cubeA_xy = xy + sampler.sample(radius, 100)
        cubeB_xy = xy + sampler.sample(radius, 100, verbose=False)
        cubeC_xy = xy + sampler.sample(radius, 100, verbose=False)
        for cube, cxy in ((self.cubeA, cubeA_xy), (self.cubeB, cubeB_xy), (self.cubeC, cubeC_xy)):
            xyz[:, :2] = cxy
            cube.set_pose(Pose.create_from_pq(xyz.clone(), U.random_quaternions(b, device=dev, lock_x=True, lock_y=True, lock_z=False)))

    # ---- stack_pyramid.py:147-189
    def _pair_ok(self, offset, cube, on_top: bool):
        flag = torch.linalg.norm(offset[..., :2], axis=1) <= torch.linalg.norm(2 * self.cube_half_size[:2]) + 0.005
        if on_top:
            flag = flag & (torch.abs(offset[..., 2]) > 0.02)
        return flag & cube.is_static(lin_thresh=1e-2, ang_thresh=0.5) & (~self.agent.is_grasping(cube))

    def evaluate(self):
        pA, pB, pC = self.cubeA.pose.p, self.cubeB.pose.p, self.cubeC.pose.p
        success = self._pair_ok(pA - pB, self.cubeA, False) & self._pair_ok(pB - pC, self.cubeC, True) & self._pair_ok(pA - pC, self.cubeC, True)
        return {"success": success}

    # ---- stack_pyramid.py:191-207
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            pA, pB, pC, tcp = self.cubeA.pose.p, self.cubeB.pose.p, self.cubeC.pose.p, self.agent.tcp.pose.p
            obs.update(cubeA_pose=self.cubeA.pose.raw_pose, cubeB_pose=self.cubeB.pose.raw_pose, cubeC_pose=self.cubeC.pose.raw_pose,
                       tcp_to_cubeA_pos=pA - tcp, tcp_to_cubeB_pos=pB - tcp, tcp_to_cubeC_pos=pC - tcp,
                       cubeA_to_cubeB_pos=pB - pA, cubeB_to_cubeC_pos=pC - pB, cubeA_to_cubeC_pos=pC - pA)
        return obs
