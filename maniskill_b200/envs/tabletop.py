"""Shared part of the tabletop Panda tasks: the robot choice (`robot_uids` = "panda" | "panda_wristcam", the SUPPORTED_ROBOTS of these
tasks that ship with assets here), loading it at the table, and `TableSceneBuilder.initialize`
(mani_skill/utils/scene_builder/table/scene_builder.py:68-127): table pose, rest configuration + Gaussian noise, gripper open, robot
base pose."""
from __future__ import annotations

import numpy as np
import torch

from ..agents import Panda
from ..model import pose7
from ..scenes import PANDA_REST_QPOS, SQRT_HALF, TABLE_HEIGHT, panda_articulation
from ..structs import Pose
from .base_env import BaseEnv

# table/scene_builder.py:73-84 ("panda") and :104-108 ("panda_wristcam": the last arm joint is turned the other way)
REST_QPOS = {"panda": PANDA_REST_QPOS,
             "panda_wristcam": np.array([0.0, np.pi / 8, 0, -np.pi * 5 / 8, 0, np.pi * 3 / 4, -np.pi / 4, 0.04, 0.04])}
ROBOT_ASSET = {"panda": "panda_v2", "panda_wristcam": "panda_v3"}


class PandaTabletopEnv(BaseEnv):
    SUPPORTED_ROBOTS = ("panda", "panda_wristcam")
    default_robot_uids = "panda"

    def __init__(self, *args, robot_uids=None, robot_init_qpos_noise=0.02, **kwargs):
        robot_uids = self.default_robot_uids if robot_uids is None else robot_uids
        if robot_uids not in self.SUPPORTED_ROBOTS:
            raise NotImplementedError(f"{type(self).__name__} on b200sim ships the robots {self.SUPPORTED_ROBOTS}, not '{robot_uids}'")
        self.robot_uids = robot_uids
        self.robot_init_qpos_noise = robot_init_qpos_noise
        if robot_uids != self.default_robot_uids:
            kwargs.setdefault("fused", False)   # fused control-step kernels are set up for a task's default robot
        super().__init__(*args, **kwargs)

    def _load_agent_desc(self):
        """`super()._load_agent(options, sapien.Pose(p=[-0.615, 0, 0]))` of the tabletop tasks."""
        art = panda_articulation(self.robot_uids, ROBOT_ASSET[self.robot_uids], (-0.615, 0, 0))
        for j in Panda.arm_joint_names:      # the drive gains follow the control mode the env is made with
            art.drive[j] = Panda.drive_gains(self._control_mode_arg)
        self.scene_desc.add_articulation(art)

    def _make_agent(self) -> Panda:
        return Panda(self.scene, self.robot_uids)

    def _robot_sensor_configs(self):
        """agents/robots/panda/panda_wristcam.py:19-32: the camera on the hand of the wrist-camera Panda."""
        if self.robot_uids != "panda_wristcam":
            return []
        return [dict(uid="hand_camera", pose=pose7(), width=128, height=128, fov=np.pi / 2, near=0.01, far=100.0, mount=(self.robot_uids, "camera_link"))]

    def _initialize_table_scene(self, env_idx: torch.Tensor):
        """table/scene_builder.py:68-127."""
        b = len(env_idx)
        dev = self.device
        self.table.set_pose(Pose.create(pose7([-0.12, 0, -TABLE_HEIGHT], [SQRT_HALF, 0, 0, SQRT_HALF]), dev))
        rest = REST_QPOS[self.robot_uids]
        if self._enhanced_determinism:   # each sub-scene draws from its own stream (scene_builder.py:85-90)
            qpos = self._batched_episode_rng[env_idx.cpu().numpy()].normal(0, self.robot_init_qpos_noise, len(rest)) + rest
        else:
            qpos = self._episode_rng.normal(0, self.robot_init_qpos_noise, (b, len(rest))) + rest
        qpos[:, -2:] = 0.04
        self.agent.reset(torch.tensor(qpos, dtype=torch.float32, device=dev))
        self.agent.robot.set_pose(Pose.create(pose7([-0.615, 0, 0]), dev))
