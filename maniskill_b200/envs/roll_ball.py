"""RollBall-v1 -- mirror of mani_skill/envs/tasks/tabletop/roll_ball.py:20-193 on the b200sim backend.

Table scene with the Panda moved to the side of the table (facing -y), a 3.5 cm ball in front of it and a non-colliding goal disc at the far
end: hit the ball so that it rolls into the goal.  The dense reward carries per-env state (`reached_status`, cleared on reset).
State observation 9 + 9 + 7 (tcp) + 3 + 7 + 3 + 3 + 3 = 44.  Task logic on the torch path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import utils as U
from .. import building as actors
from ..model import SHAPE_BOX, ActorRec, ShapeRec, pose7
from ..scenes import add_table_scene
from ..structs import Pose
from .tabletop import PandaTabletopEnv


class RollBallEnv(PandaTabletopEnv):
    max_episode_steps = 80  # @register_env("RollBall-v1", max_episode_steps=80)
    SUPPORTED_ROBOTS = ("panda",)  # roll_ball.py:36
    goal_radius = 0.1
    ball_radius = 0.035

    # ---- roll_ball.py:65-93
    def _load_scene_desc(self):
        add_table_scene(self.scene_desc)
        actors.build_sphere(self.scene_desc, radius=self.ball_radius, color=[0, 0.2, 0.8, 1], name="ball", initial_pose=actors.Pose(p=[0, 0, 0.1]))
        self.scene_desc.add_actor(ActorRec("goal_region", "kinematic",
                                           [ShapeRec(SHAPE_BOX, pose7(), np.array([1e-5, self.goal_radius, self.goal_radius]), color=(0.9, 0.1, 0.1, 1), collide=False)],
                                           pose7([0, 0, 0.1])))

    def _after_build(self):
        self.agent = self._make_agent()
        self.table = self.scene.actors["table-workspace"]
        self.ball = self.scene.actors["ball"]
        self.goal_region = self.scene.actors["goal_region"]
        self.reached_status = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)

    # ---- roll_ball.py:55-58
    def _sensor_configs(self):
        return [dict(uid="base_camera", pose=U.look_at([-0.1, 0.9, 0.3], [0.0, 0.0, 0.0]), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=None)] + self._robot_sensor_configs()

    # ---- roll_ball.py:60-63
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([-0.6, 1.3, 0.8], [0.0, 0.13, 0.0]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    # ---- table/scene_builder.py:68-103 + roll_ball.py:95-128
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        self._initialize_table_scene(env_idx)
        root_q = np.array([0.7071, 0, 0, -0.7072])  # the reference's literals (roll_ball.py:101), normalised as the simulator does
        self.agent.robot.set_pose(Pose.create(pose7([-0.1, 1.0, 0], root_q / np.linalg.norm(root_q)), dev))
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, 0] = (torch.rand((b,), device=dev) * 2 - 1) * 0.3 - 0.1
        xyz[:, 1] = torch.rand((b,), device=dev) * 0.2 + 0.5
        xyz[:, 2] = self.ball_radius
        self.ball.set_pose(Pose.create_from_pq(xyz, device=dev))
        xyz_goal = torch.zeros((b, 3), device=dev)
        xyz_goal[:, 0] = (torch.rand((b,), device=dev) * 2 - 1) * 0.3 - 0.1
        xyz_goal[:, 1] = torch.rand((b,), device=dev) * 0.2 - 1.0 + self.goal_radius
        xyz_goal[:, 2] = 1e-3
        q = torch.tensor(U.euler2quat(0, np.pi / 2, 0), dtype=torch.float32, device=dev)
        self.goal_region.set_pose(Pose.create_from_pq(xyz_goal, q[None].expand(b, 4), device=dev))
        self.reached_status[env_idx] = 0.0

    # ---- roll_ball.py:130-141
    def evaluate(self):
        return {"success": torch.linalg.norm(self.ball.pose.p[..., :2] - self.goal_region.pose.p[..., :2], axis=1) < self.goal_radius}

    # ---- roll_ball.py:143-156
    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            ball_p, goal_p = self.ball.pose.p, self.goal_region.pose.p
            obs.update(goal_pos=goal_p, ball_pose=self.ball.pose.raw_pose, ball_vel=self.ball.linear_velocity,
                       tcp_to_ball_pos=ball_p - self.agent.tcp.pose.p, ball_to_goal_pos=goal_p - ball_p)
        return obs

    # ---- roll_ball.py:158-189: reach the point 5 cm behind the ball on the goal-ball line, then the ball-to-goal distance counts
    def compute_dense_reward(self, obs, action, info):
        ball_p, goal_p = self.ball.pose.p, self.goal_region.pose.p
        unit_vec = ball_p - goal_p
        unit_vec = unit_vec / torch.linalg.norm(unit_vec, axis=1, keepdim=True)
        hit_p = ball_p + unit_vec * (self.ball_radius + 0.05)
        tcp_to_hit_dist = torch.linalg.norm(hit_p - self.agent.tcp.pose.p, axis=1)
        self.reached_status = torch.where(tcp_to_hit_dist < 0.04, 1.0, self.reached_status)
        reaching_reward = 1 - torch.tanh(2 * tcp_to_hit_dist)
        reached_reward = 1 - torch.tanh(torch.linalg.norm(ball_p[..., :2] - goal_p[..., :2], axis=1))
        reward = 20 * reached_reward * self.reached_status + reaching_reward * (1 - self.reached_status) + self.reached_status
        return torch.where(info["success"], 30.0, reward)

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 30.0
