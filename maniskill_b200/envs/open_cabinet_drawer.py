"""OpenCabinetDrawer-v1 -- mirror of mani_skill/envs/tasks/mobile_manipulation/open_cabinet_drawer.py:28-366.

The PartNet-Mobility cabinet assets are a network download that is not available; the cabinet here
is a STAND-IN built through the same path the real assets take (an articulation with a fixed root, prismatic drawer joints,
box collisions, a `handle` shape per drawer, friction 1, collision bit 29 on every link): a 0.8 x 0.5 x 0.9 m carcass with two
drawers.  Which drawer is the target is drawn per env from `_batched_episode_rng` like the reference (:131-132, :176-181),
so the per-env `handle_link` view (Link.merge in the reference) is exercised.  Robot: Fetch (29 links, 15 dof, velocity-driven
base), ground plane with the wheel / cabinet collision bits (:119-126).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import utils as U
from ..agents import Fetch
from ..model import SHAPE_PLANE, SHAPE_SPHERE, ActorRec, ArticulationRec, ShapeRec, load_robot, pose7
from ..structs import Link, Pose
from .base_env import BaseEnv

FETCH_WHEELS_COLLISION_BIT = 30
FETCH_BASE_COLLISION_BIT = 31
CABINET_COLLISION_BIT = 29

FETCH_INIT_QPOS = {"torso_lift_joint": 0.0, "head_pan_joint": 0.0, "head_tilt_joint": 0.0, "shoulder_pan_joint": 0.0,
                   "shoulder_lift_joint": -np.pi / 4, "upperarm_roll_joint": 0.0, "elbow_flex_joint": np.pi / 4, "forearm_roll_joint": 0.0,
                   "wrist_flex_joint": np.pi / 3, "wrist_roll_joint": 0.0, "l_gripper_finger_joint": 0.015, "r_gripper_finger_joint": 0.015}


def _box(name, p, half):
    return dict(type="box", p=list(p), q=[1, 0, 0, 0], half_size=list(half))


def standin_cabinet():
    """Two-drawer cabinet in the baked-robot format (tools/bake_assets.py output).  z = 0 is the cabinet bottom."""
    W, D, H, T = 0.8, 0.5, 0.9, 0.02  # width (y), depth (x), height (z), panel thickness
    fixed = dict(name="", type="fixed", p=[0, 0, 0], q=[1, 0, 0, 0], axis=[1, 0, 0], lower=0, upper=0, effort=0, damping=0, friction=0)
    carcass = [_box("back", (D / 2 - T / 2, 0, H / 2), (T / 2, W / 2, H / 2)), _box("left", (0, W / 2 - T / 2, H / 2), (D / 2, T / 2, H / 2)),
               _box("right", (0, -W / 2 + T / 2, H / 2), (D / 2, T / 2, H / 2)), _box("top", (0, 0, H - T / 2), (D / 2, W / 2, T / 2)),
               _box("bottom", (0, 0, T / 2), (D / 2, W / 2, T / 2))]
    links = [dict(name="base", parent=-1, mass=20.0, com=[0, 0, H / 2], inertia=[1, 1, 1, 0, 0, 0], collisions=carcass, joint=fixed)]
    dh = (H - 3 * T) / 2
    for i in range(2):
        zc = T + dh / 2 + i * (dh + T)
        front = _box("front", (-D / 2 + T / 2, 0, 0), (T / 2, W / 2 - T - 0.005, dh / 2 - 0.005))
        tray = _box("tray", (0.0, 0, -dh / 2 + T), (D / 2 - T, W / 2 - 2 * T, T / 2))
        handle = _box("handle", (-D / 2 - 0.03, 0, 0), (0.015, 0.08, 0.012))
        links.append(dict(name=f"drawer_{i}", parent=0, mass=3.0, com=[0, 0, 0], inertia=[0.06, 0.05, 0.1, 0, 0, 0], collisions=[front, tray, handle],
                          joint=dict(name=f"drawer_{i}_joint", type="prismatic", p=[0, 0, zc], q=[1, 0, 0, 0], axis=[-1, 0, 0], lower=0.0, upper=0.35,
                                     effort=0, damping=0, friction=0)))
    return dict(name="cabinet_standin", links=links, disabled_collision_pairs=[]), np.array([-D / 2 - 0.03, 0.0, 0.0])


class _PerEnvLink(Link):
    """Link.merge of the reference (link.py:99-126): a view whose body row differs per sub-scene."""

    def __init__(self, scene, name, rows: torch.Tensor):
        self.scene, self.name, self.row = scene, name, -1
        self._idx = torch.arange(scene.num_envs, device=scene.device) * scene.world.n_rows + rows


class _JointOfPerEnvLink:
    """`handle_link.joint` of the reference (ArticulationJoint.merge, open_cabinet_drawer.py:180-182, :308-311): the parent joint of a link
    that differs per sub-scene; `.qpos` is that joint's position in every sub-scene."""

    def __init__(self, env):
        self._env = env

    @property
    def qpos(self):
        return self._env._target_joint_qpos()


class OpenCabinetDrawerEnv(BaseEnv):
    graph_epilogue = True   # evaluate / observation / reward: pure tensor code, captured into a CUDA graph (base_env._GraphedEpilogue)
    max_episode_steps = 100
    min_open_frac = 0.75

    def __init__(self, *args, robot_uids="fetch", robot_init_qpos_noise=0.02, **kwargs):
        if robot_uids != "fetch":
            raise NotImplementedError("OpenCabinetDrawer-v1 supports robot_uids='fetch' only (SUPPORTED_ROBOTS)")
        self.robot_uids = robot_uids
        kwargs.setdefault("fused", False)
        cfg = dict(kwargs.pop("sim_config", None) or {})
        cfg.setdefault("max_contacts", 32)
        super().__init__(*args, sim_config=cfg, **kwargs)

    def _load_agent_desc(self):
        robot = load_robot("fetch")
        drive = {n: (1e3, 1e2, 100.0) for n in Fetch.arm_joint_names + Fetch.gripper_joint_names + Fetch.body_joint_names}
        drive.update({n: (0.0, 1000.0, 500.0) for n in Fetch.base_joint_names})  # PDBaseForwardVelControllerConfig(damping=1000, force_limit=500)
        art = ArticulationRec("fetch", robot, pose7([1, 0, 0]), link_mu={"l_gripper_finger_link": 2.0, "r_gripper_finger_link": 2.0},
                              disable_gravity=True, drive=drive)
        art.link_patch = {"l_gripper_finger_link": 0.1, "r_gripper_finger_link": 0.1}
        wheels = (1, 1, 1 << FETCH_WHEELS_COLLISION_BIT, 0)
        art.link_groups = {"l_wheel_link": wheels, "r_wheel_link": wheels, "base_link": (1, 1, 1 << FETCH_BASE_COLLISION_BIT, 0)}
        self.scene_desc.add_articulation(art)

    def _load_scene_desc(self):
        ground_groups = (1, 1, (1 << FETCH_WHEELS_COLLISION_BIT) | (1 << CABINET_COLLISION_BIT), 0)
        self.scene_desc.add_actor(ActorRec("ground", "static", [ShapeRec(SHAPE_PLANE, pose7([0, 0, 0], [0.7071068, 0, -0.7071068, 0]),
                                                                         color=(0.45, 0.45, 0.45, 1.0), groups=ground_groups)], pose7()))
        robot, self._handle_local = standin_cabinet()
        cab = ArticulationRec("cabinet", robot, pose7([0, 0, 0]), link_mu={L["name"]: 1.0 for L in robot["links"]}, disable_gravity=False)
        cab.link_groups = {L["name"]: (1, 1, 1 << CABINET_COLLISION_BIT, 0) for L in robot["links"]}
        self.scene_desc.add_articulation(cab)
        self._batched_episode_rng.choice(25)   # the reference draws the cabinet model first (:131, 25 ids in info_cabinet_drawer_train.json): same generator state afterwards
        self._link_ids = self._batched_episode_rng.randint(0, 2**31)  # which drawer is the target, per env (:132)
        self.scene_desc.add_actor(ActorRec("handle_link_goal", "kinematic",
                                           [ShapeRec(SHAPE_SPHERE, pose7(), np.array([0.02, 0, 0]), color=(0, 1, 0, 1), collide=False)], pose7(), hidden=True))

    def _after_build(self):
        dev = self.device
        self.agent = Fetch(self.scene, "fetch")
        self.cabinet = self.scene.articulations["cabinet"]
        self.handle_link_goal = self.scene.actors["handle_link_goal"]
        drawers = [n for n in self.cabinet.links_map if n.startswith("drawer_")]
        pick = np.asarray(self._link_ids) % len(drawers)
        rows = torch.tensor([self.cabinet.links_map[drawers[i]].row for i in pick], device=dev)
        self.handle_link = _PerEnvLink(self.scene, "handle_link", rows)
        self.handle_link.joint = _JointOfPerEnvLink(self)
        self._target_dof = torch.tensor([self.cabinet.dof_names.index(drawers[i] + "_joint") for i in pick], device=dev)
        self.handle_link_pos = torch.tensor(self._handle_local, dtype=torch.float32, device=dev)[None].expand(self.num_envs, 3)
        ar = torch.arange(self.num_envs, device=dev)
        qlim = self.cabinet.get_qlimits()[ar, self._target_dof]
        self.target_qpos = qlim[:, 0] + (qlim[:, 1] - qlim[:, 0]) * self.min_open_frac
        self.cabinet_zs = torch.zeros(self.num_envs, device=dev)

    def _target_joint_qpos(self):
        return self.cabinet.qpos[torch.arange(self.num_envs, device=self.device), self._target_dof]

    def handle_link_positions(self, env_idx=None):
        pose = self.handle_link.pose
        p = pose.p + U.quat_apply(pose.q, self.handle_link_pos)
        return p if env_idx is None else p[env_idx]

    # ---- :232-292
    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        b = len(env_idx)
        dev = self.device
        xyz = torch.zeros((b, 3), device=dev)
        xyz[:, 2] = self.cabinet_zs[env_idx]
        self.cabinet.set_pose(Pose.create_from_pq(p=xyz, device=dev))
        names = self.agent.robot.dof_names
        qpos = torch.zeros((b, len(names)), device=dev)
        for n, v in FETCH_INIT_QPOS.items():
            qpos[:, names.index(n)] = v
        dist = torch.rand((b,), device=dev) * 0.2 + 1.6
        theta = torch.rand((b,), device=dev) * (0.2 * torch.pi) + 0.9 * torch.pi
        qpos[:, names.index("root_x_axis_joint")] = torch.cos(theta) * dist
        qpos[:, names.index("root_y_axis_joint")] = torch.sin(theta) * dist
        noise_ori = torch.rand((b,), device=dev) * (0.1 * torch.pi) - 0.05 * torch.pi
        qpos[:, names.index("root_z_rotation_joint")] = (theta - torch.pi) + noise_ori
        self.agent.reset(qpos)
        self.agent.robot.set_pose(Pose.create(pose7(), dev))
        qlim = self.cabinet.get_qlimits()
        self.cabinet.set_qpos(qlim[env_idx, :, 0])
        self.cabinet.set_qvel(torch.zeros((b, self.cabinet.dof), device=dev))
        # the reference settles the cabinet with one extra px.step() on the GPU backend (:284-288)
        self.scene._gpu_apply_all()
        self.scene.world.update_kinematics()
        self.scene.step(1, 0)
        self.scene._gpu_fetch_all()
        self.handle_link_goal.set_pose(Pose.create_from_pq(p=self.handle_link_positions(env_idx), device=dev))

    # ---- :294-305
    def _after_control_step(self):
        self.handle_link_goal.set_pose(Pose.create_from_pq(p=self.handle_link_positions(), device=self.device))
        self.scene._gpu_apply_all()

    # ---- open_cabinet_drawer.py:102-107
    def _human_render_camera_configs(self):
        return [dict(uid="render_camera", pose=U.look_at([-1.8, -1.3, 1.8], [-0.3, 0.5, 0]), width=512, height=512, fov=1, near=0.01, far=100.0, mount=None)]

    def evaluate(self):
        open_enough = self._target_joint_qpos() >= self.target_qpos
        handle_link_pos = self.handle_link_positions()
        link_is_static = (torch.linalg.norm(self.handle_link.angular_velocity, axis=1) <= 1) & (torch.linalg.norm(self.handle_link.linear_velocity, axis=1) <= 0.1)
        return {"success": open_enough & link_is_static, "handle_link_pos": handle_link_pos, "open_enough": open_enough}

    def _get_obs_extra(self, info: dict):
        obs = dict(tcp_pose=self.agent.tcp.pose.raw_pose)
        if "state" in self.obs_mode:
            obs.update(tcp_to_handle_pos=info["handle_link_pos"] - self.agent.tcp.pose.p, target_link_qpos=self._target_joint_qpos(),
                       target_handle_pos=info["handle_link_pos"])
        return obs

    def compute_dense_reward(self, obs, action, info):
        tcp_to_handle_dist = torch.linalg.norm(self.agent.tcp.pose.p - info["handle_link_pos"], axis=1)
        reaching_reward = 1 - torch.tanh(5 * tcp_to_handle_dist)
        amount_to_open_left = torch.div(self.target_qpos - self._target_joint_qpos(), self.target_qpos)
        open_reward = 2 * (1 - amount_to_open_left)
        reaching_reward = torch.where(amount_to_open_left < 0.999, 2.0, reaching_reward)  # masked assignment without the nonzero() sync
        open_reward = torch.where(info["open_enough"], 3.0, open_reward)
        reward = reaching_reward + open_reward
        reward = torch.where(info["success"], 5.0, reward)
        return reward

    def compute_normalized_dense_reward(self, obs, action, info):
        return self.compute_dense_reward(obs=obs, action=action, info=info) / 5.0
