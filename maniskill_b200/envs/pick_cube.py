"""PickCube-v1 -- mirror of mani_skill/envs/tasks/tabletop/pick_cube.py:33-191 on the b200sim backend.

Scene, randomisation (same torch.rand call order under the same seed), observation layout (42-dim state), success
test and dense reward are restated line by line from the reference task; the scene prototype comes from
maniskill_b200/scenes.py.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import utils as U
from ..scenes import add_table_scene
from ..model import SHAPE_BOX, SHAPE_SPHERE, ActorRec, ShapeRec, pose7
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class PickCubeEnv(PandaTabletopEnv):
    max_episode_steps = 50  # @register_env("PickCube-v1", max_episode_steps=50)
    goal_thresh = 0.025
    cube_half_size = 0.02
    cube_spawn_half_size = 0.1
    cube_spawn_center = (0, 0)
    max_goal_height = 0.3
    sensor_cam_eye_pos = [0.3, 0, 0.6]
    sensor_cam_target_pos = [-0.1, 0, 0.1]

    # ---- pick_cube.py:79-104
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        self.scene_desc.add_actor(ActorRec("cube", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([self.cube_half_size] * 3), color=(1, 0, 0, 1))],
                                           pose7([0, 0, self.cube_half_size])))
        self.scene_desc.add_actor(ActorRec("goal_site", "kinematic",
                                           [ShapeRec(SHAPE_SPHERE, pose7(), np.array([self.goal_thresh, 0, 0]), color=(0, 1, 0, 1), collide=False)],
                                           pose7(), hidden=True))

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.cube = self.scene.actors["cube"]
        self.goal_site = self.scene.actors["goal_site"]
        self._hidden_objects.append(self.goal_site)

    def _setup_fused_step(self):
        """Fused controller + physics + evaluate/reward/obs (include/b200sim.h b2s_pick_task_*), pd_joint_delta_pos / pd_joint_pos."""
        if self.agent.control_mode not in ("pd_joint_delta_pos", "pd_joint_pos"):
            return None
        nd = self.agent.robot.dof
        dof_action = -np.ones(nd, dtype=np.int32)
        use_delta, normalize = np.zeros(nd, dtype=np.int32), np.zeros(nd, dtype=np.int32)
        low, high = np.zeros(nd, dtype=np.float32), np.zeros(nd, dtype=np.float32)
        col = 0
        for name, c in self.agent.controller.controllers.items():
            joints = c.active_joint_indices.cpu().numpy()
            if name == "gripper":  # mimic controller: one action column drives both fingers
                for j in joints:
                    dof_action[j], use_delta[j], normalize[j] = col, 0, 1
                    low[j], high[j] = float(c.action_low[0]), float(c.action_high[0])
                col += 1
            else:
                for k, j in enumerate(joints):
                    dof_action[j], use_delta[j], normalize[j] = col + k, int(c.use_delta), int(c.normalize_action)
                    low[j], high[j] = float(c.action_low[k]), float(c.action_high[k])
                col += len(joints)
        rows = dict(tcp=self.agent.tcp.row, obj=self.cube.row, goal=self.goal_site.row, lfinger=self.agent.finger1_link.row, rfinger=self.agent.finger2_link.row)
        w = self.scene.world
        handle = w.create_pick_task(dof_action, use_delta, normalize, low, high, col, rows, self.goal_thresh, nd - 2, 0,
                                    normalized_reward=self._reward_mode == "normalized_dense")
        if self._reward_mode not in ("normalized_dense", "dense"):
            return None
        N = self.num_envs
        fused = dict(handle=handle, obs=torch.zeros((N, 2 * nd + 24), dtype=torch.float32, device=self.device),
                     reward=torch.zeros(N, dtype=torch.float32, device=self.device), flags=torch.zeros((N, 6), dtype=torch.bool, device=self.device))
        # episode initialisation for the device-side auto-reset (b2s_pick_task_autoreset): what `_initialize_episode` below draws, as
        # parameters; only when this class's own initialisation is in effect (subclasses that override it keep the python reset)
        if type(self)._initialize_episode is PickCubeEnv._initialize_episode and not self._enhanced_determinism:
            from .tabletop import REST_QPOS
            w.set_pick_reset(handle, self.cube_spawn_half_size, self.cube_spawn_center, self.cube_half_size, self.max_goal_height,
                             self.robot_init_qpos_noise, REST_QPOS[self.robot_uids], self.cube.fb_index, self.goal_site.fb_index)
            fused.update(autoreset=True, final_obs=torch.zeros((N, 2 * nd + 24), dtype=torch.float32, device=self.device),
                         done=torch.zeros(N, dtype=torch.bool, device=self.device))
        return fused

    # ---- pick_cube.py:66-71
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at(self.sensor_cam_eye_pos, self.sensor_cam_target_pos), width=128, height=128,
                     fov=np.pi / 2, near=0.01, far=100.0, mount=None)] + self._robot_sensor_configs()

    # ---- pick_cube.py:73-78
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:68-103 + pick_cube.py:106-130
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, :2] = torch.rand((b, 2), device=dev) * self.cube_spawn_half_size * 2 - self.cube_spawn_half_size
        xyz[:, 0] += self.cube_spawn_center[0]
        xyz[:, 1] += self.cube_spawn_center[1]
        xyz[:, 2] = self.cube_half_size
        qs = U.random_quaternions(b, device=dev, lock_x=True, lock_y=True)
        self.cube.set_pose(Pose.create_from_pq(xyz, qs))
        goal_xyz = torch.zeros((b, 3), device=dev)
        goal_xyz[:, :2] = torch.rand((b, 2), device=dev) * self.cube_spawn_half_size * 2 - self.cube_spawn_half_size
        goal_xyz[:, 0] += self.cube_spawn_center[0]
        goal_xyz[:, 1] += self.cube_spawn_center[1]
        goal_xyz[:, 2] = torch.rand((b,), device=dev) * self.max_goal_height + xyz[:, 2]
        self.goal_site.set_pose(Pose.create_from_pq(goal_xyz, device=dev))

    # ---- pick_cube.py:132-145
    def _get_obs_extra(self, info: dict):
        obs = dict(is_grasped=info["is_grasped"], tcp_pose=self.agent.tcp_pose.raw_pose, goal_pos=self.goal_site.pose.p)
        if "state" in self.obs_mode:
            obs.update(obj_pose=self.cube.pose.raw_pose, tcp_to_obj_pos=self.cube.pose.p - self.agent.tcp_pose.p,
                       obj_to_goal_pos=self.goal_site.pose.p - self.cube.pose.p)
        return obs

    def _obs_from_fused(self, vec, info):
        """agent + extra of `get_obs` as views of the fused state vector [qpos | qvel | is_grasped, tcp_pose(7), goal_pos(3), obj_pose(7),
        tcp_to_obj_pos(3), obj_to_goal_pos(3)] (the order `_get_obs_state_dict` flattens to; pick_epilogue_kernel writes it)."""
        nd = self.agent.robot.dof
        o = 2 * nd
        extra = dict(is_grasped=info["is_grasped"], tcp_pose=vec[:, o + 1:o + 8], goal_pos=vec[:, o + 8:o + 11])
        if "state" in self.obs_mode:
            extra.update(obj_pose=vec[:, o + 11:o + 18], tcp_to_obj_pos=vec[:, o + 18:o + 21], obj_to_goal_pos=vec[:, o + 21:o + 24])
        return dict(agent=dict(qpos=vec[:, :nd], qvel=vec[:, nd:o]), extra=extra)

    # ---- pick_cube.py:147-159
    def evaluate(self):
        is_obj_placed = torch.linalg.norm(self.goal_site.pose.p - self.cube.pose.p, axis=1) <= self.goal_thresh
        is_grasped = self.agent.is_grasping(self.cube)
        is_robot_static = self.agent.is_static(0.2)
        return {"success": is_obj_placed & is_robot_static, "is_obj_placed": is_obj_placed,
                "is_robot_static": is_robot_static, "is_grasped": is_grasped}

    # ---- pick_cube.py:161-191
    def compute_dense_reward(self, obs, action, info):
        tcp_to_obj_dist = torch.linalg.norm(self.cube.pose.p - self.agent.tcp_pose.p, axis=1)
        reward = 1 - torch.tanh(5 * tcp_to_obj_dist)
        is_grasped = info["is_grasped"]
        reward = reward + is_grasped
        obj_to_goal_dist = torch.linalg.norm(self.goal_site.pose.p - self.cube.pose.p, axis=1)
        place_reward = 1 - torch.tanh(5 * obj_to_goal_dist)
        reward = reward + place_reward * is_grasped
        qvel = self.agent.robot.get_qvel()[..., :-2]
        static_reward = 1 - torch.tanh(5 * torch.linalg.norm(qvel, axis=1))
        reward = reward + static_reward * info["is_obj_placed"]
        reward = torch.where(info["success"], 5.0, reward)  # masked assignment without the nonzero() sync
        return reward

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 5
