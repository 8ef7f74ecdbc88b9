#!/usr/bin/env python
"""Generates tests/golden/host_golden.npz by running the REFERENCE's own Python functions (imported from the read-only
checkout at /root/reference) on seeded inputs.  The reference package cannot be imported as a whole here (sapien,
gymnasium, ... are not installed), so the needed files are loaded one by one with stub modules standing in for the absent
third-party packages; only pure torch/numpy code paths of the reference execute.

    python tests/golden/make_golden.py          (run in the build container; the .npz is committed)

Covered reference functions (file:line):
  mani_skill/utils/geometry/rotation_conversions.py  quaternion_raw_multiply, quaternion_apply, quaternion_to_matrix,
                                                     matrix_to_quaternion, euler_angles_to_matrix, matrix_to_euler_angles
  mani_skill/envs/utils/randomization/pose.py:13-34  random_quaternions
  mani_skill/utils/gym_utils.py:104-108              clip_and_scale_action
  mani_skill/utils/common.py:195-262, 300-304        flatten_state_dict, compute_angle_between
  mani_skill/utils/sapien_utils.py:317-366           look_at
  mani_skill/utils/structs/pose.py                   Pose.__mul__, Pose.inv, Pose.to_transformation_matrix
  mani_skill/envs/tasks/tabletop/pick_cube.py:132-191  _get_obs_extra, evaluate (success logic), compute_dense_reward
  mani_skill/agents/robots/panda/panda.py:237-269    is_grasping, is_static
  mani_skill/envs/tasks/tabletop/peg_insertion_side.py:250-360  peg_head_pose / box_hole_pose / goal_pose, has_peg_inserted,
                                                     evaluate, _get_obs_extra, compute_dense_reward
  mani_skill/envs/tasks/tabletop/push_cube.py:179-241  evaluate, _get_obs_extra, compute_dense_reward
  mani_skill/envs/tasks/tabletop/pull_cube.py:105-152  evaluate, _get_obs_extra, compute_dense_reward
  mani_skill/envs/tasks/tabletop/stack_cube.py:115-200  evaluate, _get_obs_extra, compute_dense_reward
  mani_skill/envs/tasks/tabletop/lift_peg_upright.py:88-137, poke_cube.py:126-276, roll_ball.py:130-189  the same three functions
  mani_skill/utils/structs/render_camera.py:77-155     get_extrinsic_matrix / get_model_matrix (GPU branch, mounted camera)
  mani_skill/agents/controllers/pd_ee_pose.py:85-99,229-263, utils/kinematics.py:197-260  EE controllers: action scaling, target pose, GPU IK step
  mani_skill/utils/wrappers/record.py:356-756          RecordEpisode.reset / step / flush_trajectory (h5py replaced by a dict-backed fake)
  mani_skill/envs/utils/observations/observations.py:16-68  sensor_data_to_pointcloud (two cameras)
  mani_skill/utils/wrappers/flatten.py:42-77            FlattenRGBDObservationWrapper.observation (two cameras, three settings)
  mani_skill/envs/sapien_env.py:980-1016, envs/utils/randomization/batched_rng.py  seed derivation of reset(seed=...), per-sub-scene streams
  mani_skill/utils/visualization/misc.py:54-115, sensors/camera.py:256-294  tile_images, camera_observations_to_images
  mani_skill/envs/utils/randomization/samplers.py:13-108  UniformPlacementSampler (fixed global seed)
  mani_skill/vector/wrappers/gymnasium.py:96-176     ManiSkillVectorEnv.reset / step: episode metrics, auto-reset bookkeeping
  mani_skill/agents/controllers/pd_joint_pos.py:77-101,207-228  PDJointPosController.set_action (delta / target-delta / absolute),
                                                     PDJointPosMimicController.set_action; base_controller.py:125-173 action clipping
  mani_skill/envs/tasks/mobile_manipulation/open_cabinet_drawer.py:221-358  handle_link_positions, evaluate, _get_obs_extra,
                                                     compute_dense_reward
"""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.npz")


def stub(name, **attrs):
    m = MagicMock(name=name)
    m.__name__ = name
    m.__path__ = []
    m.__all__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


def load(modname, rel):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    parent, _, child = modname.rpartition(".")
    spec.loader.exec_module(mod)
    if parent in sys.modules:
        setattr(sys.modules[parent], child, mod)
    return mod


def main():
    class FakeSapienPose:  # only used in isinstance checks
        pass

    for n in ["sapien", "sapien.physx", "sapien.render", "sapien.wrapper", "sapien.wrapper.urdf_loader", "sapien.utils", "gymnasium",
              "gymnasium.spaces", "transforms3d", "transforms3d.euler", "transforms3d.quaternions"]:
        stub(n)
    sys.modules["sapien"].Pose = FakeSapienPose
    sys.modules["gymnasium"].__version__ = "0.29.1"
    for p in ["mani_skill", "mani_skill.utils", "mani_skill.utils.geometry", "mani_skill.utils.structs", "mani_skill.envs",
              "mani_skill.envs.utils", "mani_skill.envs.tasks", "mani_skill.envs.tasks.tabletop", "mani_skill.agents",
              "mani_skill.agents.robots", "mani_skill.agents.robots.panda", "mani_skill.vector", "mani_skill.vector.wrappers"]:
        pkg(p)
    sys.modules["mani_skill"].PACKAGE_ASSET_DIR = "/nonexistent"
    stub("mani_skill.utils.logging_utils")
    stub("mani_skill.vector.wrappers.gymnasium")
    stub("mani_skill.render", SAPIEN_RENDER_SYSTEM="3.0")
    rc = load("mani_skill.utils.geometry.rotation_conversions", "mani_skill/utils/geometry/rotation_conversions.py")
    load("mani_skill.utils.structs.types", "mani_skill/utils/structs/types.py")
    common = load("mani_skill.utils.common", "mani_skill/utils/common.py")
    gym_utils = load("mani_skill.utils.gym_utils", "mani_skill/utils/gym_utils.py")
    pose_mod = load("mani_skill.utils.structs.pose", "mani_skill/utils/structs/pose.py")
    sapien_utils = load("mani_skill.utils.sapien_utils", "mani_skill/utils/sapien_utils.py")
    rnd = pkg("mani_skill.envs.utils.randomization")
    rpose = load("mani_skill.envs.utils.randomization.pose", "mani_skill/envs/utils/randomization/pose.py")
    rnd.random_quaternions = rpose.random_quaternions
    Pose = pose_mod.Pose
    G = {}
    g = torch.Generator().manual_seed(20240922)
    # ---- rotations
    qa = torch.nn.functional.normalize(torch.randn(32, 4, generator=g), dim=-1)
    qb = torch.nn.functional.normalize(torch.randn(32, 4, generator=g), dim=-1)
    v = torch.randn(32, 3, generator=g)
    G["rot_qa"], G["rot_qb"], G["rot_v"] = qa, qb, v
    G["rot_qmul"] = rc.quaternion_raw_multiply(qa, qb)
    G["rot_qapply"] = rc.quaternion_apply(qa, v)
    G["rot_q2m"] = rc.quaternion_to_matrix(qa)
    G["rot_m2q"] = rc.matrix_to_quaternion(rc.quaternion_to_matrix(qa))
    eul = torch.rand(32, 3, generator=g) * 6.28
    G["rot_euler"] = eul
    G["rot_euler_xyz_m"] = rc.euler_angles_to_matrix(eul, "XYZ")
    # ---- random_quaternions under a fixed global seed (CPU generator)
    torch.manual_seed(777)
    G["randq_lockxy"] = rpose.random_quaternions(16, lock_x=True, lock_y=True)
    torch.manual_seed(778)
    G["randq_free"] = rpose.random_quaternions(16)
    # ---- clip_and_scale_action
    a = torch.randn(8, 7, generator=g) * 1.5
    low, high = torch.full((7,), -0.1), torch.full((7,), 0.1)
    G["cs_action"], G["cs_low"], G["cs_high"] = a, low, high
    G["cs_out"] = gym_utils.clip_and_scale_action(a, low, high)
    G["cs_out_grip"] = gym_utils.clip_and_scale_action(a[:, :1], torch.tensor([-0.01]), torch.tensor([0.04]))
    # ---- flatten_state_dict / compute_angle_between
    d = dict(agent=dict(qpos=torch.randn(4, 9, generator=g), qvel=torch.randn(4, 9, generator=g)),
             extra=dict(is_grasped=torch.tensor([True, False, True, False]), tcp_pose=torch.randn(4, 7, generator=g), goal_pos=torch.randn(4, 3, generator=g)))
    G["fl_qpos"], G["fl_qvel"], G["fl_isg"], G["fl_tcp"], G["fl_goal"] = d["agent"]["qpos"], d["agent"]["qvel"], d["extra"]["is_grasped"], d["extra"]["tcp_pose"], d["extra"]["goal_pos"]
    G["fl_out"] = common.flatten_state_dict(d, use_torch=True)
    x1, x2 = torch.randn(16, 3, generator=g), torch.randn(16, 3, generator=g)
    x2[3] = 0
    G["ang_x1"], G["ang_x2"] = x1, x2
    G["ang_out"] = common.compute_angle_between(x1, x2)
    # ---- look_at (PickCube sensor + human cameras, PegInsertion base camera)
    for name, eye, tgt in [("pick_sensor", [0.3, 0, 0.6], [-0.1, 0, 0.1]), ("pick_human", [0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), ("peg_sensor", [0, -0.3, 0.2], [0, 0, 0.1])]:
        G["lookat_" + name] = sapien_utils.look_at(eye, tgt).raw_pose[0]
    # ---- Pose algebra
    pa = Pose.create_from_pq(torch.randn(8, 3, generator=g), qa[:8])
    pb = Pose.create_from_pq(torch.randn(8, 3, generator=g), qb[:8])
    G["pose_a"], G["pose_b"] = pa.raw_pose, pb.raw_pose
    G["pose_mul"] = (pa * pb).raw_pose
    G["pose_inv"] = pa.inv().raw_pose
    G["pose_mat"] = pa.to_transformation_matrix()
    # ---- PickCube task logic on synthetic states
    stub("mani_skill.agents.robots", SO100=object, Fetch=object, Panda=object, WidowXAI=object, XArm6Robotiq=object)
    sapien_env = stub("mani_skill.envs.sapien_env")
    sapien_env.BaseEnv = type("BaseEnv", (), {})
    stub("mani_skill.sensors")
    stub("mani_skill.sensors.camera")
    stub("mani_skill.utils.building")
    reg = stub("mani_skill.utils.registration")
    reg.register_env = lambda *a, **k: (lambda cls: cls)
    stub("mani_skill.utils.scene_builder")
    stub("mani_skill.utils.scene_builder.table")
    load("mani_skill.envs.tasks.tabletop.pick_cube_cfgs", "mani_skill/envs/tasks/tabletop/pick_cube_cfgs.py")
    pc = load("mani_skill.envs.tasks.tabletop.pick_cube", "mani_skill/envs/tasks/tabletop/pick_cube.py")
    n = 12
    cube_p = torch.randn(n, 3, generator=g) * 0.1
    goal_p = cube_p + torch.randn(n, 3, generator=g) * 0.03
    goal_p[:3] = cube_p[:3] + 0.001
    tcp_raw = torch.hstack([cube_p + torch.randn(n, 3, generator=g) * 0.05, qa[:n]])
    cube_raw = torch.hstack([cube_p, qb[:n]])
    qvel = torch.randn(n, 9, generator=g) * 0.3
    qvel[:2] *= 0.01
    is_grasped = torch.tensor([True, False] * (n // 2))
    is_static = torch.max(torch.abs(qvel[:, :-2]), 1)[0] <= 0.2
    fake = SimpleNamespace(
        cube=SimpleNamespace(pose=Pose.create(cube_raw)), goal_site=SimpleNamespace(pose=Pose.create_from_pq(goal_p)),
        agent=SimpleNamespace(tcp_pose=Pose.create(tcp_raw), robot=SimpleNamespace(get_qvel=lambda: qvel),
                              is_grasping=lambda obj: is_grasped, is_static=lambda thr: is_static),
        goal_thresh=0.025, robot_uids="panda", obs_mode="state")
    info = pc.PickCubeEnv.evaluate(fake)
    fake.compute_dense_reward = lambda obs, action, info: pc.PickCubeEnv.compute_dense_reward(fake, obs, action, info)
    G["pc_cube"], G["pc_goal"], G["pc_tcp"], G["pc_qvel"], G["pc_is_grasped"] = cube_raw, goal_p, tcp_raw, qvel, is_grasped
    G["pc_success"], G["pc_is_obj_placed"], G["pc_is_robot_static"] = info["success"], info["is_obj_placed"], info["is_robot_static"]
    G["pc_reward"] = pc.PickCubeEnv.compute_dense_reward(fake, None, None, info)
    G["pc_reward_norm"] = pc.PickCubeEnv.compute_normalized_dense_reward(fake, None, None, info)
    extra = pc.PickCubeEnv._get_obs_extra(fake, info)
    G["pc_extra_flat"] = common.flatten_state_dict(extra, use_torch=True)
    # ---- PegInsertionSide task logic on synthetic states (peg near / inside / far from the hole, grasped or not)
    class SapienPose:  # stands in for sapien.Pose in `tgt_gripper_pose * sapien.Pose([-0.06, 0, 0])`
        def __init__(self, p=(0, 0, 0), q=(1, 0, 0, 0)):
            self.p, self.q = np.asarray(p, dtype=np.float32), np.asarray(q, dtype=np.float32)

    sys.modules["sapien"].Pose = SapienPose
    stub("mani_skill.utils.building.actors")
    sys.modules["mani_skill.agents.robots.panda"].PandaWristCam = object
    stub("mani_skill.envs.scene")
    sys.modules["mani_skill.utils.structs"].Pose = Pose
    sys.modules["mani_skill.utils.structs"].Actor = object
    sys.modules["mani_skill.envs.utils"].randomization = sys.modules["mani_skill.envs.utils.randomization"]
    sys.modules["mani_skill.utils"].common = common
    sys.modules["mani_skill.utils"].sapien_utils = sapien_utils
    peg_mod = load("mani_skill.envs.tasks.tabletop.peg_insertion_side", "mani_skill/envs/tasks/tabletop/peg_insertion_side.py")
    PE = peg_mod.PegInsertionSideEnv
    m = 16
    lengths = 0.085 + 0.04 * torch.rand(m, generator=g)
    radii = 0.015 + 0.01 * torch.rand(m, generator=g)
    peg_half = torch.stack([lengths, radii, radii], 1)
    box_raw = torch.hstack([torch.randn(m, 3, generator=g) * 0.1, torch.nn.functional.normalize(torch.randn(m, 4, generator=g), dim=-1)])
    hole_off = torch.hstack([torch.zeros(m, 1), 0.02 * torch.randn(m, 2, generator=g)])
    box_hole_offsets = Pose.create_from_pq(p=hole_off)
    peg_head_offsets = Pose.create_from_pq(p=torch.hstack([lengths[:, None], torch.zeros(m, 2)]))
    hole_radii = radii + 0.003
    goal = Pose.create(box_raw) * box_hole_offsets * peg_head_offsets.inv()
    # envs 0-3: peg exactly at the goal (inserted); 4-7: slightly off (pre-inserted / close); the rest: anywhere
    pert = torch.randn(m, 3, generator=g) * 0.1
    pert[:4] = 0
    pert[4:8] = torch.randn(4, 3, generator=g) * 0.004
    peg_raw = goal.raw_pose.clone()
    peg_raw[:, :3] += pert
    peg_raw[8:, 3:] = torch.nn.functional.normalize(torch.randn(m - 8, 4, generator=g), dim=-1)
    tcp_raw2 = torch.hstack([peg_raw[:, :3] + torch.randn(m, 3, generator=g) * 0.05, qa[:m]])
    peg_grasped = torch.tensor([True, True, False, True] * (m // 4))
    fake_peg = SimpleNamespace(
        peg=SimpleNamespace(pose=Pose.create(peg_raw)), box=SimpleNamespace(pose=Pose.create(box_raw)), peg_head_offsets=peg_head_offsets,
        box_hole_offsets=box_hole_offsets, box_hole_radii=hole_radii, peg_half_sizes=peg_half,
        agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(tcp_raw2)), is_grasping=lambda obj, max_angle=None: peg_grasped),
        obs_mode_struct=SimpleNamespace(use_state=True))
    fake_peg.peg_head_pose = PE.peg_head_pose.fget(fake_peg)
    fake_peg.box_hole_pose = PE.box_hole_pose.fget(fake_peg)
    fake_peg.goal_pose = PE.goal_pose.fget(fake_peg)
    fake_peg.has_peg_inserted = lambda: PE.has_peg_inserted(fake_peg)
    pinfo = PE.evaluate(fake_peg)
    G["peg_peg"], G["peg_box"], G["peg_tcp"], G["peg_half"], G["peg_hole_off"], G["peg_hole_radii"], G["peg_grasped"] = \
        peg_raw, box_raw, tcp_raw2, peg_half, hole_off, hole_radii, peg_grasped
    G["peg_head_pose"], G["peg_box_hole_pose"], G["peg_goal_pose"] = fake_peg.peg_head_pose.raw_pose, fake_peg.box_hole_pose.raw_pose, fake_peg.goal_pose.raw_pose
    G["peg_success"], G["peg_head_at_hole"] = pinfo["success"], pinfo["peg_head_pos_at_hole"]
    G["peg_reward"] = PE.compute_dense_reward(fake_peg, None, None, pinfo)
    G["peg_extra_flat"] = common.flatten_state_dict(PE._get_obs_extra(fake_peg, pinfo), use_torch=True)
    # ---- OpenCabinetDrawer task logic on synthetic states (drawer closed / partly open / open enough, handle moving or still)
    for n_ in ["trimesh", "mani_skill.utils.building.articulations", "mani_skill.utils.building.ground", "mani_skill.utils.io_utils"]:
        stub(n_)
    pkg("mani_skill.envs.tasks.mobile_manipulation")
    import pathlib
    sys.modules["mani_skill"].PACKAGE_ASSET_DIR = pathlib.Path("/nonexistent")
    stub("mani_skill.utils.geometry.bounding_cylinder")
    geo = load("mani_skill.utils.geometry.geometry", "mani_skill/utils/geometry/geometry.py")
    sys.modules["mani_skill.utils.structs"].Articulation = object
    sys.modules["mani_skill.utils.structs"].Link = object
    cab_mod = load("mani_skill.envs.tasks.mobile_manipulation.open_cabinet_drawer", "mani_skill/envs/tasks/mobile_manipulation/open_cabinet_drawer.py")
    CE = cab_mod.OpenCabinetDrawerEnv
    m = 12
    target_qpos = 0.2 + 0.2 * torch.rand(m, generator=g)
    joint_qpos = target_qpos * torch.tensor([0.0, 0.0005, 0.3, 0.95, 1.0, 1.05, 1.2, 0.5, 1.1, 0.0, 1.3, 0.999])
    handle_pose = Pose.create_from_pq(torch.randn(m, 3, generator=g) * 0.3, torch.nn.functional.normalize(torch.randn(m, 4, generator=g), dim=-1))
    handle_local = torch.randn(m, 3, generator=g) * 0.1
    ang_v = torch.randn(m, 3, generator=g) * 0.8
    lin_v = torch.randn(m, 3, generator=g) * 0.08
    tcp3 = Pose.create_from_pq(torch.randn(m, 3, generator=g) * 0.3, qa[:m])
    fake_cab = SimpleNamespace(
        handle_link=SimpleNamespace(joint=SimpleNamespace(qpos=joint_qpos), pose=handle_pose, angular_velocity=ang_v, linear_velocity=lin_v),
        handle_link_pos=handle_local, target_qpos=target_qpos, device=torch.device("cpu"), obs_mode="state",
        agent=SimpleNamespace(tcp=SimpleNamespace(pose=tcp3)))
    fake_cab.handle_link_positions = lambda env_idx=None: CE.handle_link_positions(fake_cab, env_idx)
    cinfo = CE.evaluate(fake_cab)
    G["cab_target_qpos"], G["cab_joint_qpos"], G["cab_handle_pose"], G["cab_handle_local"], G["cab_ang_v"], G["cab_lin_v"], G["cab_tcp"] = \
        target_qpos, joint_qpos, handle_pose.raw_pose, handle_local, ang_v, lin_v, tcp3.raw_pose
    G["cab_success"], G["cab_open_enough"], G["cab_handle_link_pos"] = cinfo["success"], cinfo["open_enough"], cinfo["handle_link_pos"]
    G["cab_reward"] = CE.compute_dense_reward(fake_cab, None, None, cinfo)
    G["cab_extra_flat"] = common.flatten_state_dict(CE._get_obs_extra(fake_cab, cinfo), use_torch=True)
    # ---- Panda.is_grasping / is_static
    base_agent = stub("mani_skill.agents.base_agent")
    base_agent.BaseAgent = type("BaseAgent", (), {})
    base_agent.Keyframe = lambda **k: SimpleNamespace(**k)
    stub("mani_skill.agents.controllers")
    regs = stub("mani_skill.agents.registration")
    regs.register_agent = lambda *a, **k: (lambda cls: cls)
    stub("mani_skill.utils.structs.actor")
    sys.modules["sapien"].Pose = lambda *a, **k: None
    panda = load("mani_skill.agents.robots.panda.panda", "mani_skill/agents/robots/panda/panda.py")
    lf = torch.randn(n, 3, generator=g) * 3
    rf = torch.randn(n, 3, generator=g) * 3
    lf[:2] = 0
    f1 = Pose.create_from_pq(torch.randn(n, 3, generator=g), qa[8:8 + n])
    f2 = Pose.create_from_pq(torch.randn(n, 3, generator=g), qb[8:8 + n])
    fl1, fl2 = SimpleNamespace(pose=f1), SimpleNamespace(pose=f2)
    fake_agent = SimpleNamespace(finger1_link=fl1, finger2_link=fl2, robot=SimpleNamespace(get_qvel=lambda: qvel),
                                 scene=SimpleNamespace(get_pairwise_contact_forces=lambda a, b: lf if a is fl1 else rf))
    G["pg_lforce"], G["pg_rforce"], G["pg_f1"], G["pg_f2"] = lf, rf, f1.raw_pose, f2.raw_pose
    G["pg_is_grasping"] = panda.Panda.is_grasping(fake_agent, None)
    G["pg_is_static"] = panda.Panda.is_static(fake_agent, 0.2)
    # ---- PushCube task logic on synthetic states (appended after the older sections so that their random draws do not move)
    g3 = torch.Generator().manual_seed(4242)
    sys.modules["sapien"].Pose = FakeSapienPose  # isinstance checks in Pose.create need a class again
    sys.modules["transforms3d.euler"].euler2quat = lambda *a, **k: np.array([1.0, 0, 0, 0])
    push_mod = load("mani_skill.envs.tasks.tabletop.push_cube", "mani_skill/envs/tasks/tabletop/push_cube.py")
    PU = push_mod.PushCubeEnv
    m = 12
    obj_p = torch.hstack([torch.randn(m, 2, generator=g3) * 0.1, torch.full((m, 1), 0.02)])
    obj_p[:3, 2] += torch.tensor([0.004, 0.006, 0.02])            # lifted cubes: one inside, two outside the 5 mm band
    goal_p = torch.hstack([obj_p[:, :2] + torch.randn(m, 2, generator=g3) * 0.12, torch.full((m, 1), 1e-3)])
    goal_p[:4, :2] = obj_p[:4, :2] + 0.01
    obj_raw = torch.hstack([obj_p, torch.nn.functional.normalize(torch.randn(m, 4, generator=g3), dim=-1)])
    tcp_p = obj_p + torch.randn(m, 3, generator=g3) * 0.05
    tcp_p[4:8] = obj_p[4:8] + torch.tensor([-0.025, 0.0, 0.0]) + torch.randn(4, 3, generator=g3) * 0.002   # at the push pose
    tcp_raw3 = torch.hstack([tcp_p, torch.nn.functional.normalize(torch.randn(m, 4, generator=g3), dim=-1)])
    fake_push = SimpleNamespace(obj=SimpleNamespace(pose=Pose.create(obj_raw)), goal_region=SimpleNamespace(pose=Pose.create_from_pq(goal_p)),
                                agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(tcp_raw3))), goal_radius=0.1, cube_half_size=0.02,
                                device=torch.device("cpu"), obs_mode_struct=SimpleNamespace(use_state=True))
    uinfo = PU.evaluate(fake_push)
    G["push_obj"], G["push_goal"], G["push_tcp"] = obj_raw, goal_p, tcp_raw3
    G["push_success"] = uinfo["success"]
    G["push_reward"] = PU.compute_dense_reward(fake_push, None, None, uinfo)
    G["push_extra_flat"] = common.flatten_state_dict(PU._get_obs_extra(fake_push, uinfo), use_torch=True)
    # ---- UniformPlacementSampler under a fixed global seed (three sequential samples, tight bounds so that rejections happen)
    samp = load("mani_skill.envs.utils.randomization.samplers", "mani_skill/envs/utils/randomization/samplers.py")
    torch.manual_seed(31337)
    sp = samp.UniformPlacementSampler(bounds=[[-0.05, -0.06], [0.05, 0.06]], batch_size=16)
    G["sampler_pts"] = torch.stack([sp.sample(0.03, 100), sp.sample(0.03, 100, verbose=False), sp.sample(0.02, 100, verbose=False)])
    # ---- StackCube task logic on synthetic states
    sys.modules["mani_skill.envs.utils"].randomization = sys.modules["mani_skill.envs.utils.randomization"]
    stack_mod = load("mani_skill.envs.tasks.tabletop.stack_cube", "mani_skill/envs/tasks/tabletop/stack_cube.py")
    SC = stack_mod.StackCubeEnv
    m = 12
    B_p = torch.hstack([torch.randn(m, 2, generator=g3) * 0.1, torch.full((m, 1), 0.02)])
    A_p = B_p + torch.randn(m, 3, generator=g3) * 0.08
    A_p[:6] = B_p[:6] + torch.tensor([0.0, 0.0, 0.04]) + torch.randn(6, 3, generator=g3) * torch.tensor([0.012, 0.012, 0.003])  # (nearly) stacked
    A_raw = torch.hstack([A_p, torch.nn.functional.normalize(torch.randn(m, 4, generator=g3), dim=-1)])
    B_raw = torch.hstack([B_p, torch.nn.functional.normalize(torch.randn(m, 4, generator=g3), dim=-1)])
    A_lin = torch.randn(m, 3, generator=g3) * 0.01
    A_ang = torch.randn(m, 3, generator=g3) * 0.4
    A_lin[:4] *= 0.1
    A_ang[:4] *= 0.1
    tcp4 = torch.hstack([A_p + torch.randn(m, 3, generator=g3) * 0.04, torch.nn.functional.normalize(torch.randn(m, 4, generator=g3), dim=-1)])
    grasped4 = torch.tensor([False, True, False, False, True, False] * 2)
    qpos4 = torch.rand(m, 9, generator=g3) * 0.04
    qlim = torch.zeros(1, 9, 2)
    qlim[0, :, 1] = 0.04
    cubeA = SimpleNamespace(pose=Pose.create(A_raw), linear_velocity=A_lin, angular_velocity=A_ang,
                            is_static=lambda lin_thresh=1e-2, ang_thresh=1e-1: (torch.linalg.norm(A_lin, axis=1) <= lin_thresh) & (torch.linalg.norm(A_ang, axis=1) <= ang_thresh))
    fake_stack = SimpleNamespace(cubeA=cubeA, cubeB=SimpleNamespace(pose=Pose.create(B_raw)), cube_half_size=torch.tensor([0.02] * 3), device=torch.device("cpu"),
                                 obs_mode="state", agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(tcp4)), is_grasping=lambda obj: grasped4,
                                                                         robot=SimpleNamespace(get_qlimits=lambda: qlim, get_qpos=lambda: qpos4)))
    sinfo = SC.evaluate(fake_stack)
    G["stack_A"], G["stack_B"], G["stack_A_lin"], G["stack_A_ang"], G["stack_tcp"], G["stack_grasped"], G["stack_qpos"] = A_raw, B_raw, A_lin, A_ang, tcp4, grasped4, qpos4
    for k_ in ("is_cubeA_on_cubeB", "is_cubeA_static", "success"):
        G["stack_" + k_] = sinfo[k_]
    G["stack_reward"] = SC.compute_dense_reward(fake_stack, None, None, sinfo)
    G["stack_extra_flat"] = common.flatten_state_dict(SC._get_obs_extra(fake_stack, sinfo), use_torch=True)
    # ---- joint-space controllers: the reference's own set_action on a stand-in controller object (row a1 of SURVEY section 8)
    stub("gymnasium.vector")
    stub("gymnasium.vector.utils")
    stub("mani_skill.agents.utils")
    sys.modules["mani_skill.utils"].gym_utils = gym_utils
    sys.modules["mani_skill.utils.structs"].ArticulationJoint = object
    pkg("mani_skill.agents.controllers")
    bc = load("mani_skill.agents.controllers.base_controller", "mani_skill/agents/controllers/base_controller.py")
    pjp = load("mani_skill.agents.controllers.pd_joint_pos", "mani_skill/agents/controllers/pd_joint_pos.py")
    nenv = 6
    qpos_arm = torch.randn(nenv, 7, generator=g3) * 0.5
    acts = [torch.randn(nenv, 7, generator=g3) * 0.8 for _ in range(2)]
    sent = []

    def fake_ctrl(cls, qpos, low, high, use_delta, use_target, normalize=True, **extra):
        c = SimpleNamespace(config=SimpleNamespace(use_delta=use_delta, use_target=use_target, interpolate=False), scene=SimpleNamespace(num_envs=nenv),
                            action_space=SimpleNamespace(shape=(nenv, low.shape[0])), _normalize_action=normalize, action_space_low=low, action_space_high=high,
                            qpos=qpos, _target_qpos=qpos.clone(), _start_qpos=qpos.clone(), **extra)
        c._preprocess_action = lambda a: bc.BaseController._preprocess_action(c, a)
        c._clip_and_scale_action = lambda a: bc.BaseController._clip_and_scale_action(c, a)
        c.set_drive_targets = lambda t: sent.append(t.clone())
        return c

    low7, high7 = torch.full((7,), -0.1), torch.full((7,), 0.1)
    G["ctl_qpos_arm"], G["ctl_act0"], G["ctl_act1"] = qpos_arm, acts[0], acts[1]
    c = fake_ctrl(pjp.PDJointPosController, qpos_arm, low7, high7, True, False)
    pjp.PDJointPosController.set_action(c, acts[0])
    G["ctl_delta_target"] = sent[-1]
    c = fake_ctrl(pjp.PDJointPosController, qpos_arm, low7, high7, True, True)
    pjp.PDJointPosController.set_action(c, acts[0])
    pjp.PDJointPosController.set_action(c, acts[1])
    G["ctl_target_delta_target"] = sent[-1]
    c = fake_ctrl(pjp.PDJointPosController, qpos_arm, low7, high7, False, False, normalize=False)
    pjp.PDJointPosController.set_action(c, acts[0])
    G["ctl_abs_target"] = sent[-1]
    qpos_grip = torch.rand(nenv, 2, generator=g3) * 0.04
    act_grip = torch.randn(nenv, 1, generator=g3)
    c = fake_ctrl(pjp.PDJointPosMimicController, qpos_grip, torch.tensor([-0.01]), torch.tensor([0.04]), False, False,
                  control_joint_indices=torch.tensor([0]), mimic_joint_indices=torch.tensor([1]), mimic_control_joint_indices=torch.tensor([0]),
                  _multiplier=torch.ones(1), _offset=torch.zeros(1))
    pjp.PDJointPosMimicController.set_action(c, act_grip)
    G["ctl_qpos_grip"], G["ctl_act_grip"], G["ctl_mimic_target"] = qpos_grip, act_grip, sent[-1]
    # ---- PDBaseForwardVelController.set_action (Fetch base, pd_base_vel.py:39-73)
    stub("mani_skill.agents.controllers.pd_joint_vel", PDJointVelController=object, PDJointVelControllerConfig=object)
    pbv = load("mani_skill.agents.controllers.pd_base_vel", "mani_skill/agents/controllers/pd_base_vel.py")
    base_q = torch.hstack([torch.randn(nenv, 2, generator=g3), torch.rand(nenv, 1, generator=g3) * 6.28 - 3.14])
    base_act = torch.randn(nenv, 2, generator=g3) * 0.9
    vel_sent = []
    cb = SimpleNamespace(qpos=base_q, joints=None, active_joint_indices=None, scene=SimpleNamespace(num_envs=nenv), action_space=SimpleNamespace(shape=(nenv, 2)),
                         _normalize_action=True, action_space_low=torch.tensor([-1.0, -3.14]), action_space_high=torch.tensor([1.0, 3.14]),
                         articulation=SimpleNamespace(set_joint_drive_velocity_targets=lambda t, j, idx: vel_sent.append(t.clone())))
    cb._preprocess_action = lambda a: bc.BaseController._preprocess_action(cb, a)
    cb._clip_and_scale_action = lambda a: bc.BaseController._clip_and_scale_action(cb, a)
    pbv.PDBaseForwardVelController.set_action(cb, base_act)
    G["ctl_base_q"], G["ctl_base_act"], G["ctl_base_vel_target"] = base_q, base_act, vel_sent[-1]
    # ---- ManiSkillVectorEnv.step / reset on a scripted inner env (record_metrics, partial auto-reset)
    sys.modules["gymnasium.vector"].VectorEnv = object
    sys.modules["gymnasium"].vector = sys.modules["gymnasium.vector"]
    del sys.modules["mani_skill.vector.wrappers.gymnasium"]
    vw = load("mani_skill.vector.wrappers.gymnasium", "mani_skill/vector/wrappers/gymnasium.py")
    nv, T_ = 5, 7
    script_rew = torch.rand(T_, nv, generator=g3)
    script_succ = torch.rand(T_, nv, generator=g3) < 0.25
    script_trunc = torch.zeros(T_, nv, dtype=torch.bool)
    script_trunc[4, :] = True
    G["vec_rew"], G["vec_succ"], G["vec_trunc"] = script_rew, script_succ, script_trunc

    class ScriptedEnv:  # elapsed_steps + scripted (obs, reward, terminated, truncated, info); partial reset zeroes elapsed_steps
        def __init__(self):
            self.t = 0
            self.elapsed_steps = torch.zeros(nv, dtype=torch.int32)
            self.device = torch.device("cpu")
            self.num_envs = nv
        def step(self, a):
            self.elapsed_steps = self.elapsed_steps + 1
            t = self.t
            self.t += 1
            return (torch.full((nv, 2), float(t)), script_rew[t].clone(), script_succ[t].clone(), script_trunc[t].clone(),
                    dict(success=script_succ[t].clone(), elapsed_steps=self.elapsed_steps.clone()))
        def reset(self, seed=None, options=None):
            idx = options["env_idx"] if options and "env_idx" in options else torch.arange(nv)
            self.elapsed_steps[idx] = 0
            return torch.full((nv, 2), -1.0), dict(reset=True)

    inner = ScriptedEnv()
    wrap = SimpleNamespace(_env=inner, base_env=inner, num_envs=nv, auto_reset=True, ignore_terminations=False, record_metrics=True, device=inner.device,
                           success_once=torch.zeros(nv, dtype=torch.bool), fail_once=torch.zeros(nv, dtype=torch.bool), returns=torch.zeros(nv))
    wrap.reset = lambda seed=None, options=None: vw.ManiSkillVectorEnv.reset(wrap, seed=seed, options=options)
    for t in range(T_):
        o, r, te, tr, info = vw.ManiSkillVectorEnv.step(wrap, None)
        G[f"vec_obs_{t}"], G[f"vec_term_{t}"], G[f"vec_truncout_{t}"] = o, te, tr
        ep = info["final_info"]["episode"] if "final_info" in info else info["episode"]
        G[f"vec_has_final_{t}"] = torch.tensor("final_info" in info)
        for k_ in ("success_once", "return", "episode_len", "reward"):
            G[f"vec_{k_}_{t}"] = ep[k_]
        G[f"vec_returns_after_{t}"] = wrap.returns.clone()
    # ---- PullCube task logic (same synthetic states as PushCube, pull pose instead of push pose)
    pull_mod = load("mani_skill.envs.tasks.tabletop.pull_cube", "mani_skill/envs/tasks/tabletop/pull_cube.py")
    PL = pull_mod.PullCubeEnv
    tcp_pl = tcp_raw3.clone()
    tcp_pl[4:8, :3] = obj_p[4:8] + torch.tensor([0.03, 0.0, 0.0]) + torch.randn(4, 3, generator=g3) * 0.002   # at the pull pose
    fake_pull = SimpleNamespace(obj=fake_push.obj, goal_region=fake_push.goal_region, agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(tcp_pl))),
                                goal_radius=0.1, cube_half_size=0.02, device=torch.device("cpu"), obs_mode_struct=SimpleNamespace(use_state=True))
    linfo = PL.evaluate(fake_pull)
    G["pull_tcp"], G["pull_success"] = tcp_pl, linfo["success"]
    G["pull_reward"] = PL.compute_dense_reward(fake_pull, None, None, linfo)
    G["pull_extra_flat"] = common.flatten_state_dict(PL._get_obs_extra(fake_pull, linfo), use_torch=True)
    # ---- matrix_to_euler_angles (the IK step of the end-effector controllers, agents/controllers/utils/kinematics.py:233-236)
    g2 = torch.Generator().manual_seed(99)
    qe = torch.nn.functional.normalize(torch.randn(24, 4, generator=g2), dim=-1)
    qe[:8, 1:] *= 0.05  # small rotations, the regime the controllers use
    qe = torch.nn.functional.normalize(qe, dim=-1)
    G["eul_q"] = qe
    G["eul_xyz_from_matrix"] = rc.matrix_to_euler_angles(rc.quaternion_to_matrix(qe), "XYZ")
    # ---- LiftPegUpright / PokeCube / RollBall task logic on synthetic states (own generator: the sections above keep their draws)
    g4 = torch.Generator().manual_seed(777)
    rnd_q = lambda k: torch.nn.functional.normalize(torch.randn(k, 4, generator=g4), dim=-1)
    lift_mod = load("mani_skill.envs.tasks.tabletop.lift_peg_upright", "mani_skill/envs/tasks/tabletop/lift_peg_upright.py")
    LP = lift_mod.LiftPegUprightEnv
    m = 12
    # lying (the reset pose), upright (tipped about world y, a little wobble) and arbitrary pegs
    q_lie = torch.tensor([0.5 ** 0.5, 0.5 ** 0.5, 0.0, 0.0]).expand(m, 4)
    tip = torch.tensor([0.5 ** 0.5, 0.0, -(0.5 ** 0.5), 0.0]).expand(m, 4)
    wob = torch.nn.functional.normalize(torch.hstack([torch.ones(m, 1), torch.randn(m, 3, generator=g4) * 0.02]), dim=-1)
    q_up = rc.quaternion_multiply(wob, rc.quaternion_multiply(tip, q_lie))
    peg_q = torch.vstack([q_lie[:3], q_up[3:9], rnd_q(3)])
    peg_p = torch.hstack([torch.randn(m, 2, generator=g4) * 0.1, torch.full((m, 1), 0.025)])
    peg_p[3:9, 2] = 0.12 + torch.tensor([0.0, 0.001, -0.002, 0.004, 0.006, 0.0])
    peg_raw = torch.hstack([peg_p, peg_q])
    tcp_lp = torch.hstack([peg_p + torch.randn(m, 3, generator=g4) * 0.05, rnd_q(m)])
    grasp_lp = torch.rand(m, generator=g4) < 0.4
    peg_ns = SimpleNamespace(pose=Pose.create(peg_raw))
    fake_lp = SimpleNamespace(peg=peg_ns, peg_half_length=0.12, device=torch.device("cpu"), obs_mode_struct=SimpleNamespace(use_state=True),
                              agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(tcp_lp)), is_grasping=lambda o: grasp_lp.clone()))
    pinfo = LP.evaluate(fake_lp)
    G["lift_peg"], G["lift_tcp"], G["lift_grasp"], G["lift_success"] = peg_raw, tcp_lp, grasp_lp, pinfo["success"]
    G["lift_reward"] = LP.compute_dense_reward(fake_lp, None, None, pinfo)
    G["lift_extra_flat"] = common.flatten_state_dict(LP._get_obs_extra(fake_lp, pinfo), use_torch=True)
    # PokeCube
    poke_mod = load("mani_skill.envs.tasks.tabletop.poke_cube", "mani_skill/envs/tasks/tabletop/poke_cube.py")
    PK = poke_mod.PokeCubeEnv
    yaw_q = lambda a: torch.stack([torch.cos(a / 2), torch.zeros_like(a), torch.zeros_like(a), torch.sin(a / 2)], dim=1)
    pk_peg_yaw = torch.randn(m, generator=g4) * 0.2
    pk_cube_yaw = pk_peg_yaw + torch.randn(m, generator=g4) * 0.06     # aligned within 0.05 rad for some, not for others
    pk_peg_p = torch.hstack([torch.randn(m, 2, generator=g4) * 0.1, torch.full((m, 1), 0.025)])
    pk_cube_p = pk_peg_p + torch.hstack([0.12 + torch.rand(m, 1, generator=g4) * 0.06, torch.randn(m, 1, generator=g4) * 0.01, torch.full((m, 1), -0.005)])
    pk_goal_p = torch.hstack([pk_cube_p[:, :2] + torch.randn(m, 2, generator=g4) * 0.06, torch.full((m, 1), 1e-3)])
    pk_peg_raw, pk_cube_raw = torch.hstack([pk_peg_p, yaw_q(pk_peg_yaw)]), torch.hstack([pk_cube_p, yaw_q(pk_cube_yaw)])
    pk_tcp = torch.hstack([pk_peg_p + torch.randn(m, 3, generator=g4) * 0.02, rnd_q(m)])
    pk_tcp[:6, :3] = pk_peg_p[:6] + torch.randn(6, 3, generator=g4) * 0.002    # within the 1 cm "reached" ball
    pk_grasp = torch.rand(m, generator=g4) < 0.6
    pk_static = torch.rand(m, generator=g4) < 0.7
    pk_static[0] = False   # a placed cube with the arm still moving
    pk_qvel = torch.randn(m, 9, generator=g4) * 0.1
    fake_pk = SimpleNamespace(cube=SimpleNamespace(pose=Pose.create(pk_cube_raw)), peg=SimpleNamespace(pose=Pose.create(pk_peg_raw)),
                              goal_region=SimpleNamespace(pose=Pose.create_from_pq(pk_goal_p)), goal_radius=0.05, cube_half_size=0.02,
                              peg_head_offsets=Pose.create_from_pq(p=[0.12, 0, 0]), device=torch.device("cpu"), obs_mode_struct=SimpleNamespace(use_state=True),
                              agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(pk_tcp)), is_grasping=lambda o: pk_grasp.clone(),
                                                    is_static=lambda t: pk_static.clone(), robot=SimpleNamespace(get_qvel=lambda: pk_qvel)))
    fake_pk.peg_head_pos = PK.peg_head_pos.fget(fake_pk)
    fake_pk.peg_head_pose = PK.peg_head_pose.fget(fake_pk)
    kinfo = PK.evaluate(fake_pk)
    G["poke_peg"], G["poke_cube"], G["poke_goal"], G["poke_tcp"] = pk_peg_raw, pk_cube_raw, pk_goal_p, pk_tcp
    G["poke_grasp"], G["poke_static"], G["poke_qvel"] = pk_grasp, pk_static, pk_qvel
    for k_ in ("success", "is_cube_placed", "is_peg_cube_fit", "is_peg_grasped", "angle_diff", "head_to_cube_dist"):
        G[f"poke_{k_}"] = kinfo[k_]
    G["poke_reward"] = PK.compute_dense_reward(fake_pk, None, None, kinfo)
    G["poke_extra_flat"] = common.flatten_state_dict(PK._get_obs_extra(fake_pk, kinfo), use_torch=True)
    # RollBall (reached_status latches inside compute_dense_reward)
    roll_mod = load("mani_skill.envs.tasks.tabletop.roll_ball", "mani_skill/envs/tasks/tabletop/roll_ball.py")
    RB = roll_mod.RollBallEnv
    rb_ball = torch.hstack([torch.randn(m, 2, generator=g4) * 0.3, torch.full((m, 1), 0.035), rnd_q(m)])
    rb_goal = torch.hstack([rb_ball[:, :2] + torch.randn(m, 2, generator=g4) * 0.5, torch.full((m, 1), 1e-3)])
    rb_goal[:3, :2] = rb_ball[:3, :2] + 0.03
    unit = torch.nn.functional.normalize(rb_ball[:, :3] - rb_goal, dim=1)
    rb_tcp = torch.hstack([rb_ball[:, :3] + torch.randn(m, 3, generator=g4) * 0.2, rnd_q(m)])
    rb_tcp[3:7, :3] = rb_ball[3:7, :3] + unit[3:7] * 0.085 + torch.randn(4, 3, generator=g4) * 0.01    # at the hit point
    rb_vel = torch.randn(m, 3, generator=g4)
    rb_status0 = (torch.rand(m, generator=g4) < 0.3).float()
    fake_rb = SimpleNamespace(ball=SimpleNamespace(pose=Pose.create(rb_ball), linear_velocity=rb_vel), goal_region=SimpleNamespace(pose=Pose.create_from_pq(rb_goal)),
                              goal_radius=0.1, ball_radius=0.035, reached_status=rb_status0.clone(), device=torch.device("cpu"),
                              obs_mode_struct=SimpleNamespace(use_state=True), agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(rb_tcp))))
    binfo = RB.evaluate(fake_rb)
    G["roll_ball"], G["roll_goal"], G["roll_tcp"], G["roll_vel"], G["roll_status0"] = rb_ball, rb_goal, rb_tcp, rb_vel, rb_status0
    G["roll_success"] = binfo["success"]
    G["roll_reward"] = RB.compute_dense_reward(fake_rb, None, None, binfo)
    G["roll_status1"] = fake_rb.reached_status.clone()
    G["roll_extra_flat"] = common.flatten_state_dict(RB._get_obs_extra(fake_rb, binfo), use_torch=True)
    # ---- camera parameters (mani_skill/utils/structs/render_camera.py:77-155, GPU branch): extrinsic_cv and the GL model matrix
    stub("mani_skill.render", SAPIEN_RENDER_SYSTEM="3.0")
    for mname in ("mani_skill.utils.structs.actor", "mani_skill.utils.structs.link"):
        if mname not in sys.modules:
            stub(mname, Actor=object, Link=object)
    rcam = load("mani_skill.utils.structs.render_camera", "mani_skill/utils/structs/render_camera.py")
    mount_raw = torch.hstack([torch.randn(m, 3, generator=g4), rnd_q(m)])
    local_raw = torch.hstack([torch.randn(1, 3, generator=g4) * 0.1, rnd_q(1)])
    cam_global = Pose.create(mount_raw) * Pose.create(local_raw)
    fake_cam = SimpleNamespace(scene=SimpleNamespace(gpu_sim_enabled=True, device=torch.device("cpu")), mount=object(), global_pose=cam_global,
                               _cached_extrinsic_matrix=None, _cached_model_matrix=None)
    fake_cam.get_global_pose = lambda: fake_cam.global_pose
    G["cam_mount"], G["cam_local"] = mount_raw, local_raw
    G["cam_extrinsic_cv"] = rcam.RenderCamera.get_extrinsic_matrix(fake_cam)
    G["cam_model_gl"] = rcam.RenderCamera.get_model_matrix(fake_cam)
    # ---- end-effector controllers (pd_ee_pose.py:85-99,229-263) and the GPU IK step (utils/kinematics.py:197-260) with a scripted Jacobian
    stub("pytorch_kinematics")
    stub("lxml", etree=MagicMock())
    stub("lxml.etree")
    stub("sapien.wrapper.pinocchio_model", PinocchioModel=object)
    for mname, attrs in (("mani_skill.utils.structs.articulation", dict(Articulation=object)), ("mani_skill.utils.structs.articulation_joint", dict(ArticulationJoint=object))):
        if mname not in sys.modules:
            stub(mname, **attrs)
    kin_mod = load("mani_skill.agents.controllers.utils.kinematics", "mani_skill/agents/controllers/utils/kinematics.py")
    if not isinstance(getattr(sys.modules["mani_skill.utils"], "sapien_utils", None), types.ModuleType):
        sys.modules["mani_skill.utils"].sapien_utils = MagicMock()
    ee_mod = load("mani_skill.agents.controllers.pd_ee_pose", "mani_skill/agents/controllers/pd_ee_pose.py")
    ne = 10
    ee_act = torch.randn(ne, 6, generator=g4) * 0.8
    ee_act[:3, 3:] *= 3.0                                   # rotation parts with norm > 1 are clipped by norm
    # action_space_low/high: the bounds the normalised action is scaled to (panda.py: pos_lower/upper = -+0.1, rot_lower/upper = -+0.1)
    ee_self = SimpleNamespace(action_space_low=torch.tensor([-0.1] * 6), action_space_high=torch.tensor([0.1] * 6),
                              config=SimpleNamespace(rot_lower=-0.1, use_delta=True, frame="root_translation:root_aligned_body_rotation"))
    ee_scaled = ee_mod.PDEEPoseController._clip_and_scale_action(ee_self, ee_act)
    prev_raw = torch.hstack([torch.randn(ne, 3, generator=g4) * 0.3, rnd_q(ne)])
    G["ee_act"], G["ee_scaled"], G["ee_prev"] = ee_act, ee_scaled, prev_raw
    G["ee_target_pose"] = ee_mod.PDEEPoseController.compute_target_pose(ee_self, Pose.create(prev_raw), ee_scaled).raw_pose
    pos_self = SimpleNamespace(config=SimpleNamespace(use_delta=True, frame="root_translation"))
    G["ee_target_pos_only"] = ee_mod.PDEEPosController.compute_target_pose(pos_self, Pose.create(prev_raw), ee_scaled[:, :3]).raw_pose
    J = torch.randn(ne, 6, 7, generator=g4)
    q0 = torch.randn(ne, 9, generator=g4)
    cur_raw = torch.hstack([prev_raw[:, :3] + torch.randn(ne, 3, generator=g4) * 0.02, rc.quaternion_multiply(
        torch.nn.functional.normalize(torch.hstack([torch.ones(ne, 1), torch.randn(ne, 3, generator=g4) * 0.05]), dim=-1), prev_raw[:, 3:])])
    kin_self = SimpleNamespace(use_gpu_ik=True, active_ancestor_joint_idxs=torch.arange(7), qmask=torch.ones(7, dtype=torch.bool), device=torch.device("cpu"),
                               pk_chain=SimpleNamespace(jacobian=lambda q: J))
    G["ee_J"], G["ee_q0"], G["ee_cur"] = J, q0, cur_raw
    for sname, cfg in (("lm", dict(type="levenberg_marquardt", alpha=1.0)), ("pinv", dict(type="pseudo_inverse", alpha=0.5))):
        # virtual-target path: a target Pose + the current pose; direct path: the scaled action as a 6-vector delta
        G[f"ee_ik_target_{sname}"] = kin_mod.Kinematics.compute_ik(kin_self, Pose.create(G["ee_target_pose"]), q0, is_delta_pose=False,
                                                                   current_pose=Pose.create(cur_raw), solver_config=cfg)
        G[f"ee_ik_delta_{sname}"] = kin_mod.Kinematics.compute_ik(kin_self, ee_scaled.clone(), q0, is_delta_pose=True, current_pose=Pose.create(cur_raw),
                                                                  solver_config=cfg)
    # ---- RecordEpisode (mani_skill/utils/wrappers/record.py:356-756) on a scripted 3-sub-scene env; h5py replaced by a dict-backed fake
    class FakeGroup:
        def __init__(self, store, path):
            self.store, self.path = store, path
        def create_group(self, name, track_order=True):
            return FakeGroup(self.store, f"{self.path}/{name}" if self.path else name)
        def create_dataset(self, key, data=None, dtype=None, **kw):
            self.store[f"{self.path}/{key}"] = np.array(data, dtype=dtype)

    class FakeWrapper:
        def __init__(self, env):
            self.env = env
        def reset(self, *a, **k):
            return self.env.reset(*a, **k)
        def step(self, a):
            return self.env.step(a)

    sys.modules["gymnasium"].Wrapper = FakeWrapper
    stub("h5py", File=object, Group=object)
    stub("mani_skill.utils.io_utils", dump_json=lambda path, data, indent=2: None)
    stub("mani_skill.utils.logging_utils")
    stub("mani_skill.utils.visualization")
    stub("mani_skill.utils.visualization.misc")
    stub("mani_skill.utils.wrappers", CPUGymWrapper=type("CPUGymWrapper", (), {}))
    sys.modules["mani_skill"].get_commit_info = lambda: {}
    if not hasattr(sys.modules["mani_skill.utils"].sapien_utils, "is_state_dict_consistent") or isinstance(sys.modules["mani_skill.utils"].sapien_utils, MagicMock):
        sys.modules["mani_skill.utils"].sapien_utils = SimpleNamespace(is_state_dict_consistent=lambda sd: True)
    rec_mod = load("mani_skill.utils.wrappers.record", "mani_skill/utils/wrappers/record.py")
    nr, Tr = 3, 8
    rs_obs = torch.randn(Tr + 4, nr, 4, generator=g4)          # what the scripted env returns, frame by frame (resets consume frames too)
    rs_rew = torch.rand(Tr, nr, generator=g4)
    rs_succ = torch.rand(Tr, nr, generator=g4) < 0.4
    rs_state = torch.randn(Tr + 4, nr, 5, generator=g4)
    rs_act = torch.randn(Tr, nr, 2, generator=g4)
    G["rec_obs"], G["rec_rew"], G["rec_succ"], G["rec_state"], G["rec_act"] = rs_obs, rs_rew, rs_succ, rs_state, rs_act

    class ScriptedRecEnv:
        def __init__(self):
            self.k, self.t, self.num_envs, self.control_mode = 0, 0, nr, "pd_joint_delta_pos"
            self._episode_seed = np.array([5, 6, 7])
            self.unwrapped = self
            self.cur = torch.zeros(nr, 5)
        def get_wrapper_attr(self, name):
            return SimpleNamespace(sample=lambda: np.zeros(2, dtype=np.float32))
        def get_state_dict(self):
            return dict(actors=dict(cube=self.cur.clone()), articulations=dict(panda=self.cur.clone() * 2))
        def reset(self, seed=None, options=None):
            idx = torch.arange(nr) if not options or "env_idx" not in options else torch.as_tensor(options["env_idx"])
            self.cur[idx] = rs_state[self.k][idx]
            obs = rs_obs[self.k].clone()
            self.k += 1
            return obs, dict(reconfigure=False)
        def step(self, a):
            self.cur = rs_state[self.k].clone()
            obs = rs_obs[self.k].clone()
            self.k += 1
            t = self.t
            self.t += 1
            return obs, rs_rew[t].clone(), rs_succ[t].clone(), torch.zeros(nr, dtype=torch.bool), dict(success=rs_succ[t].clone())

    store = {}
    rec = rec_mod.RecordEpisode.__new__(rec_mod.RecordEpisode)
    rec.env = ScriptedRecEnv()
    rec.__dict__.update(_h5_file=FakeGroup(store, ""), _json_data=dict(episodes=[]), _json_path="x.json", _trajectory_buffer=None, save_on_reset=True,
                        save_trajectory=True, record_env_state=True, record_reward=True, _episode_id=-1, _elapsed_record_steps=0, _save_video=False,
                        save_video_trigger=None, cpu_wrapped_env=False, _already_warned_about_state_dict_inconsistency=False, last_reset_kwargs={})
    # script: reset, 3 steps, partial reset of sub-scene 1, 2 steps, partial reset of sub-scenes 0 and 2, 2 steps, full reset, 1 step, flush
    rec.reset(seed=5)
    for t in range(3):
        rec.step(rs_act[t])
    rec.reset(options=dict(env_idx=torch.tensor([1])))
    for t in range(3, 5):
        rec.step(rs_act[t])
    rec.reset(options=dict(env_idx=torch.tensor([0, 2])))
    for t in range(5, 7):
        rec.step(rs_act[t])
    rec.reset()
    rec.step(rs_act[7])
    rec.flush_trajectory()
    for k_, v_ in store.items():
        G["recout/" + k_] = v_
    G["rec_episode_steps"] = np.array([e["elapsed_steps"] for e in rec._json_data["episodes"]])
    G["rec_episode_seed"] = np.array([e["episode_seed"] for e in rec._json_data["episodes"]])
    G["rec_episode_success"] = np.array([e["success"] for e in rec._json_data["episodes"]])
    # ---- sensor_data_to_pointcloud (mani_skill/envs/utils/observations/observations.py:16-68): two cameras, synthetic targets
    FakeCamera = type("Camera", (), {})
    stub("mani_skill.sensors")
    stub("mani_skill.sensors.base_sensor", BaseSensor=object, BaseSensorConfig=object)
    stub("mani_skill.sensors.camera", Camera=FakeCamera)
    obs_mod = load("mani_skill.envs.utils.observations.observations", "mani_skill/envs/utils/observations/observations.py")
    npc, Hh, Ww = 2, 4, 5
    pc_obs = dict(sensor_data={}, sensor_param={})
    for ci, uid in enumerate(("base_camera", "hand_camera")):
        pos = (torch.randn(npc, Hh, Ww, 3, generator=g4) * 400).to(torch.int16)
        seg = (torch.rand(npc, Hh, Ww, 1, generator=g4) * 4).to(torch.int16)          # id 0 = background
        rgb = (torch.rand(npc, Hh, Ww, 3, generator=g4) * 255).to(torch.uint8)
        c2w = torch.eye(4).repeat(npc, 1, 1)
        c2w[:, :3, :3] = rc.quaternion_to_matrix(rnd_q(npc))
        c2w[:, :3, 3] = torch.randn(npc, 3, generator=g4)
        pc_obs["sensor_data"][uid] = dict(rgb=rgb, position=pos, segmentation=seg)
        pc_obs["sensor_param"][uid] = dict(cam2world_gl=c2w)
        G[f"pcd_in_{ci}_position"], G[f"pcd_in_{ci}_segmentation"], G[f"pcd_in_{ci}_rgb"], G[f"pcd_in_{ci}_cam2world"] = pos, seg, rgb, c2w
    pc_out = obs_mod.sensor_data_to_pointcloud(pc_obs, {"base_camera": FakeCamera(), "hand_camera": FakeCamera()})
    assert pc_out["sensor_data"] == {}
    for k_ in ("xyzw", "rgb", "segmentation"):
        G[f"pcd_out_{k_}"] = pc_out["pointcloud"][k_]
    # ---- FlattenRGBDObservationWrapper.observation (mani_skill/utils/wrappers/flatten.py:42-77), two cameras
    class FakeObsWrapper:
        def __init__(self, env):
            self.env = env
    sys.modules["gymnasium"].ObservationWrapper = FakeObsWrapper
    sys.modules["gymnasium"].ActionWrapper = FakeObsWrapper
    stub("gymnasium.spaces")
    stub("gymnasium.spaces.utils")
    stub("gymnasium.vector.utils", batch_space=None)
    flat_mod = load("mani_skill.utils.wrappers.flatten", "mani_skill/utils/wrappers/flatten.py")
    nf = 3
    def fl_obs():
        gg = torch.Generator().manual_seed(31)
        cam = lambda: dict(rgb=(torch.rand(nf, 4, 5, 3, generator=gg) * 255).to(torch.uint8), depth=(torch.rand(nf, 4, 5, 1, generator=gg) * 2000).to(torch.int16))
        return dict(agent=dict(qpos=torch.randn(nf, 9, generator=gg), qvel=torch.randn(nf, 9, generator=gg)),
                    extra=dict(is_grasped=torch.rand(nf, generator=gg) < 0.5, tcp_pose=torch.randn(nf, 7, generator=gg)),
                    sensor_param=dict(base_camera=dict(), hand_camera=dict()), sensor_data=dict(base_camera=cam(), hand_camera=cam()))
    src = fl_obs()
    G["flat_in_state_parts"] = torch.hstack([src["agent"]["qpos"], src["agent"]["qvel"], src["extra"]["is_grasped"][:, None].float(), src["extra"]["tcp_pose"]])
    for ci, uid in enumerate(("base_camera", "hand_camera")):
        G[f"flat_in_rgb_{ci}"], G[f"flat_in_depth_{ci}"] = src["sensor_data"][uid]["rgb"], src["sensor_data"][uid]["depth"]
    for tag, kw in (("sep", dict(include_rgb=True, include_depth=True, sep_depth=True, include_state=True)),
                    ("merged", dict(include_rgb=True, include_depth=True, sep_depth=False, include_state=True)),
                    ("rgbonly", dict(include_rgb=True, include_depth=False, sep_depth=True, include_state=False))):
        fw = SimpleNamespace(base_env=SimpleNamespace(device=torch.device("cpu")), **kw)
        out = flat_mod.FlattenRGBDObservationWrapper.observation(fw, fl_obs())
        G[f"flat_keys_{tag}"] = np.array(sorted(out.keys()))
        for k_, v_ in out.items():
            G[f"flat_{tag}_{k_}"] = v_
    # ---- BaseEnv._set_main_rng / _set_episode_rng (mani_skill/envs/sapien_env.py:980-1016) + BatchedRNG (batched_rng.py): seed derivation
    brng = load("mani_skill.envs.utils.randomization.batched_rng", "mani_skill/envs/utils/randomization/batched_rng.py")
    sys.modules["gymnasium"].Env = type("Env", (), {})
    for mname in ("mani_skill.agents", "mani_skill.envs.scene", "mani_skill.envs.utils.observations", "mani_skill.envs.utils.system",
                  "mani_skill.envs.utils.system.backend", "mani_skill.sensors.depth_camera", "mani_skill.utils.structs", "mani_skill.utils.structs.types",
                  "mani_skill.utils.visualization.misc", "mani_skill.render.utils", "mani_skill.utils.tree", "dacite", "mani_skill.examples",
                  "mani_skill.examples.real2sim_3d_assets", "sapien.utils", "sapien.utils.viewer", "sapien.utils.viewer.control_window", "sapien.render",
                  "mani_skill.envs.utils", "mani_skill.envs.utils.randomization"):
        if mname not in sys.modules or not isinstance(sys.modules[mname], (MagicMock, types.ModuleType)):
            stub(mname)
    import ast
    for node in ast.walk(ast.parse(open(os.path.join(REF, "mani_skill/envs/sapien_env.py")).read())):   # every imported name must exist
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in ("mani_skill", "sapien", "gymnasium", "dacite"):
            if node.module not in sys.modules:
                stub(node.module)
            for a in node.names:
                if not hasattr(sys.modules[node.module], a.name):
                    setattr(sys.modules[node.module], a.name, MagicMock(name=a.name))
        elif isinstance(node, ast.Import):
            for a in node.names:
                if a.name.split(".")[0] in ("sapien", "gymnasium", "dacite") and a.name not in sys.modules:
                    stub(a.name)
    saved_base_env_stub = sys.modules.get("mani_skill.envs.sapien_env")
    try:
        se = load("mani_skill.envs.sapien_env_real", "mani_skill/envs/sapien_env.py")
        RB_ = se.BaseEnv
    finally:
        if saved_base_env_stub is not None:
            sys.modules["mani_skill.envs.sapien_env"] = saved_base_env_stub
    se.BatchedRNG = brng.BatchedRNG
    nv = 4
    fe = SimpleNamespace(num_envs=nv, _main_seed=None, _batched_rng_backend="numpy:random_state", _enhanced_determinism=False,
                         _episode_seed=np.zeros(nv, dtype=np.int64), _batched_episode_rng=None)
    trace = []
    def snap(tag):
        trace.append(tag)
        G[f"rng_{tag}_main_seed"] = np.asarray(fe._main_seed, dtype=np.int64)
        G[f"rng_{tag}_episode_seed"] = np.asarray(fe._episode_seed, dtype=np.int64).copy()
        G[f"rng_{tag}_draw"] = fe._batched_episode_rng.uniform(0, 1, size=(2,))          # advances the episode streams, like a task would
        G[f"rng_{tag}_env0_normal"] = fe._episode_rng.normal(0, 0.02, (2, 3))
    # reset(seed=7): one seed fans out to every sub-scene; then partial reset without seed under enhanced determinism; then a list of seeds
    RB_._set_main_rng(fe, 7)
    RB_._set_episode_rng(fe, 7, torch.arange(nv))
    snap("a")
    RB_._set_main_rng(fe, None)                           # unseeded reset: main rng untouched, episode rng keeps running
    RB_._set_episode_rng(fe, None, torch.arange(nv))
    snap("b")
    fe._enhanced_determinism = True
    RB_._set_main_rng(fe, None)
    RB_._set_episode_rng(fe, None, torch.tensor([1, 3]))  # fresh episode seeds for the sub-scenes being reset, drawn from their main rngs
    snap("c")
    RB_._set_main_rng(fe, [11, 12, 13, 14])
    RB_._set_episode_rng(fe, [11, 12, 13, 14], torch.arange(nv))
    snap("d")
    # ---- BaseEnv.step / get_reward / compute_sparse_reward (sapien_env.py:648-700,1042-1071) for every success / fail combination
    ns = 6
    st_succ, st_fail = torch.rand(ns, generator=g4) < 0.5, torch.rand(ns, generator=g4) < 0.3
    st_dense = torch.rand(ns, generator=g4)
    G["step_succ"], G["step_fail"], G["step_dense"] = st_succ, st_fail, st_dense
    for tag, keys in (("sf", ("success", "fail")), ("s", ("success",)), ("f", ("fail",)), ("none", ())):
        for mode in ("sparse", "dense", "normalized_dense", "none"):
            if (tag, mode) == ("f", "sparse"):
                continue        # sapien_env.py:693 negates a bool tensor, which torch refuses: the reference raises here
            info0 = {k_: dict(success=st_succ, fail=st_fail)[k_].clone() for k_ in keys}
            fs = SimpleNamespace(num_envs=ns, device=torch.device("cpu"), _elapsed_steps=torch.zeros(ns, dtype=torch.int32), _reward_mode=mode,
                                 _step_action=lambda a: a, get_info=lambda: dict(info0), get_obs=lambda info, unflattened=True: dict(x=torch.ones(ns, 2)),
                                 _flatten_raw_obs=lambda o: o["x"], compute_dense_reward=lambda obs, action, info: st_dense * 5,
                                 compute_normalized_dense_reward=lambda obs, action, info: st_dense)
            fs.get_reward = lambda obs, action, info: RB_.get_reward(fs, obs, action, info)
            fs.compute_sparse_reward = lambda obs, action, info: RB_.compute_sparse_reward(fs, obs, action, info)
            o_, r_, te_, tr_, i_ = RB_.step(fs, torch.zeros(ns, 3))
            G[f"step_{tag}_{mode}_reward"] = torch.as_tensor(r_).float()
            G[f"step_{tag}_{mode}_terminated"], G[f"step_{tag}_{mode}_truncated"] = te_, tr_
            assert int(fs._elapsed_steps[0]) == 1
    # ---- tile_images (mani_skill/utils/visualization/misc.py:54-115) and camera_observations_to_images (mani_skill/sensors/camera.py:256-294)
    for mname in ("imageio", "tqdm", "PIL", "PIL.Image", "PIL.ImageDraw", "PIL.ImageFont"):
        if mname not in sys.modules:
            stub(mname)
    misc = load("mani_skill.utils.visualization.misc_real", "mani_skill/utils/visualization/misc.py")
    gi = torch.Generator().manual_seed(77)
    rnd_img = lambda *shape: (torch.rand(*shape, generator=gi) * 255).to(torch.uint8)
    tile_sets = dict(a=[rnd_img(2, 8, 8, 3), rnd_img(2, 4, 4, 3), rnd_img(2, 4, 4, 3)],                 # render_all of PickCube: big view + two small
                     b=[rnd_img(2, 4, 6, 3), rnd_img(2, 4, 6, 3), rnd_img(2, 4, 6, 3)],                 # equal sizes: one per column
                     c=[rnd_img(3, 5, 3), rnd_img(6, 4, 3), rnd_img(3, 4, 3), rnd_img(2, 5, 3)])        # unbatched, mixed
    for tag, imgs in tile_sets.items():
        for i_, im in enumerate(imgs):
            G[f"tile_{tag}_in{i_}"] = im
        G[f"tile_{tag}_out"] = misc.tile_images([im.clone() for im in imgs])
    G["tile_b_out_2rows"] = misc.tile_images([im.clone() for im in tile_sets["b"]] + [tile_sets["b"][0].clone()], nrows=2)
    import ast as _ast
    cam_src = open(os.path.join(REF, "mani_skill/sensors/camera.py")).read()
    for node in _ast.walk(_ast.parse(cam_src)):
        if isinstance(node, _ast.ImportFrom) and node.module and node.module.split(".")[0] in ("mani_skill", "sapien"):
            if node.module not in sys.modules:
                stub(node.module)
            for a in node.names:
                if not hasattr(sys.modules[node.module], a.name):
                    setattr(sys.modules[node.module], a.name, MagicMock(name=a.name))
    cam_mod = load("mani_skill.sensors.camera_real", "mani_skill/sensors/camera.py")
    co = dict(rgb=rnd_img(2, 4, 5, 3), depth=(torch.rand(2, 4, 5, 1, generator=gi) * 1500).to(torch.int16),
              segmentation=(torch.rand(2, 4, 5, 1, generator=gi) * 20).to(torch.int16), position=(torch.randn(2, 4, 5, 3, generator=gi) * 500).to(torch.int16))
    for k_, v_ in co.items():
        G[f"camimg_in_{k_}"] = v_
    for k_, v_ in cam_mod.camera_observations_to_images({k2: v2.clone() for k2, v2 in co.items()}).items():
        G[f"camimg_out_{k_}"] = v_
    # ---- PDJointVelController / PDJointPosVelController.set_action (pd_joint_vel.py:38-42, pd_joint_pos_vel.py:43-67)
    sys.modules["mani_skill.agents.controllers.pd_joint_pos"] = pjp
    pjv = load("mani_skill.agents.controllers.pd_joint_vel_real", "mani_skill/agents/controllers/pd_joint_vel.py")
    gv = torch.Generator().manual_seed(404)
    v_act = torch.randn(nenv, 7, generator=gv) * 0.9
    vel_sent2, pos_sent2 = [], []
    cv = SimpleNamespace(joints=None, active_joint_indices=None, scene=SimpleNamespace(num_envs=nenv), action_space=SimpleNamespace(shape=(nenv, 7)),
                         _normalize_action=True, action_space_low=torch.full((7,), -1.0), action_space_high=torch.full((7,), 1.0),
                         articulation=SimpleNamespace(set_joint_drive_velocity_targets=lambda t, j, idx: vel_sent2.append(t.clone())))
    cv._preprocess_action = lambda a: bc.BaseController._preprocess_action(cv, a)
    cv._clip_and_scale_action = lambda a: bc.BaseController._clip_and_scale_action(cv, a)
    pjv.PDJointVelController.set_action(cv, v_act)
    G["ctl_vel_act"], G["ctl_vel_target"] = v_act, vel_sent2[-1]
    pjpv = load("mani_skill.agents.controllers.pd_joint_pos_vel_real", "mani_skill/agents/controllers/pd_joint_pos_vel.py")
    pv_acts = [torch.randn(nenv, 14, generator=gv) * 0.8 for _ in range(2)]
    lowpv, highpv = torch.cat([torch.full((7,), -0.1), torch.full((7,), -1.0)]), torch.cat([torch.full((7,), 0.1), torch.full((7,), 1.0)])
    cpv = fake_ctrl(pjpv.PDJointPosVelController, qpos_arm, lowpv, highpv, True, True)
    cpv.action_space = SimpleNamespace(shape=(nenv, 14))
    cpv.set_drive_targets = lambda t: pos_sent2.append(t.clone())
    cpv.set_drive_velocity_targets = lambda t: vel_sent2.append(t.clone())
    for a_ in pv_acts:
        pjpv.PDJointPosVelController.set_action(cpv, a_)
    G["ctl_pv_act0"], G["ctl_pv_act1"], G["ctl_pv_pos_target"], G["ctl_pv_vel_target"] = pv_acts[0], pv_acts[1], pos_sent2[-1], vel_sent2[-1]
    # ---- PlaceSphere task logic (place_sphere.py:186-265)
    for mname in ("matplotlib", "matplotlib.pyplot"):
        if mname not in sys.modules:
            stub(mname)
    ps_mod = load("mani_skill.envs.tasks.tabletop.place_sphere", "mani_skill/envs/tasks/tabletop/place_sphere.py")
    PS = ps_mod.PlaceSphereEnv
    gp = torch.Generator().manual_seed(515)
    mps = 12
    bin_p = torch.hstack([torch.rand(mps, 1, generator=gp) * 0.1, torch.rand(mps, 1, generator=gp) * 0.2 - 0.1, torch.full((mps, 1), 0.0025)])
    sph_p = bin_p + torch.hstack([torch.randn(mps, 2, generator=gp) * 0.05, torch.rand(mps, 1, generator=gp) * 0.1])
    sph_p[:6] = bin_p[:6] + torch.tensor([0.0, 0.0, 0.0225]) + torch.randn(6, 3, generator=gp) * 0.002            # in the bin
    sph_raw = torch.hstack([sph_p, torch.nn.functional.normalize(torch.randn(mps, 4, generator=gp), dim=-1)])
    ps_lin, ps_ang = torch.randn(mps, 3, generator=gp) * 0.008, torch.randn(mps, 3, generator=gp) * 0.3
    ps_tcp = torch.hstack([sph_p + torch.randn(mps, 3, generator=gp) * 0.03, torch.nn.functional.normalize(torch.randn(mps, 4, generator=gp), dim=-1)])
    ps_lin[:3] *= 0.1
    ps_ang[:3] *= 0.1                                          # three resting spheres among the ones in the bin
    ps_grasp = torch.rand(mps, generator=gp) < 0.4
    ps_grasp[0], ps_grasp[1] = False, True                     # one released (success), one still held
    ps_rstatic = torch.rand(mps, generator=gp) < 0.6
    ps_qpos = torch.hstack([torch.randn(mps, 7, generator=gp), torch.rand(mps, 2, generator=gp) * 0.04])
    ps_qlim = torch.tensor([[-2.9, 2.9]] * 7 + [[0.0, 0.04]] * 2)[None].repeat(mps, 1, 1)
    obj_ns = SimpleNamespace(pose=Pose.create(sph_raw), linear_velocity=ps_lin, angular_velocity=ps_ang)
    obj_ns.is_static = lambda lin_thresh=1e-2, ang_thresh=1e-1: torch.logical_and(torch.linalg.norm(ps_lin, axis=1) <= lin_thresh, torch.linalg.norm(ps_ang, axis=1) <= ang_thresh)
    fake_ps = SimpleNamespace(obj=obj_ns, bin=SimpleNamespace(pose=Pose.create_from_pq(bin_p)), radius=0.02, block_half_size=PS.block_half_size, device=torch.device("cpu"),
                              obs_mode="state", agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(ps_tcp)), is_grasping=lambda o: ps_grasp.clone(),
                                                                      is_static=lambda t: ps_rstatic.clone(),
                                                                      robot=SimpleNamespace(get_qlimits=lambda: ps_qlim, get_qpos=lambda: ps_qpos)))
    sinfo = PS.evaluate(fake_ps)
    G["place_sphere"], G["place_bin"], G["place_lin"], G["place_ang"], G["place_tcp"] = sph_raw, bin_p, ps_lin, ps_ang, ps_tcp
    G["place_grasp"], G["place_rstatic"], G["place_qpos"] = ps_grasp, ps_rstatic, ps_qpos
    for k_ in ("is_obj_grasped", "is_obj_on_bin", "is_obj_static", "success"):
        G[f"place_{k_}"] = sinfo[k_]
    G["place_reward"] = PS.compute_dense_reward(fake_ps, None, None, sinfo)
    G["place_extra_flat"] = common.flatten_state_dict(PS._get_obs_extra(fake_ps, sinfo), use_torch=True)
    # ---- StackPyramid task logic (stack_pyramid.py:147-207)
    sp_mod = load("mani_skill.envs.tasks.tabletop.stack_pyramid", "mani_skill/envs/tasks/tabletop/stack_pyramid.py")
    SPy = sp_mod.StackPyramidEnv
    gq = torch.Generator().manual_seed(616)
    msp = 12
    pB = torch.hstack([torch.randn(msp, 2, generator=gq) * 0.1, torch.full((msp, 1), 0.02)])
    pA = pB + torch.hstack([torch.randn(msp, 2, generator=gq) * 0.04, torch.zeros(msp, 1)])
    pC = (pA + pB) / 2 + torch.hstack([torch.randn(msp, 2, generator=gq) * 0.03, torch.full((msp, 1), 0.04)])
    pA[:6] = pB[:6] + torch.tensor([0.0, 0.041, 0.0])
    pC[:6] = (pA[:6] + pB[:6]) / 2 + torch.tensor([0.0, 0.0, 0.04]) + torch.randn(6, 3, generator=gq) * 0.002       # pyramids
    pC[4, 2] = 0.02                                                                                                    # blue cube not on top
    rawq = lambda p_: torch.hstack([p_, torch.nn.functional.normalize(torch.randn(msp, 4, generator=gq), dim=-1)])
    rA, rB, rC = rawq(pA), rawq(pB), rawq(pC)
    stat = {n_: torch.rand(msp, generator=gq) < 0.8 for n_ in "ABC"}
    grsp = {n_: torch.rand(msp, generator=gq) < 0.15 for n_ in "ABC"}
    sp_tcp = torch.hstack([pC + torch.randn(msp, 3, generator=gq) * 0.05, torch.nn.functional.normalize(torch.randn(msp, 4, generator=gq), dim=-1)])
    cubes = {n_: SimpleNamespace(pose=Pose.create(r_), tag=n_) for n_, r_ in zip("ABC", (rA, rB, rC))}
    for n_, c_ in cubes.items():
        c_.is_static = (lambda s_: (lambda lin_thresh=1e-2, ang_thresh=0.5: s_.clone()))(stat[n_])
    fake_sp = SimpleNamespace(cubeA=cubes["A"], cubeB=cubes["B"], cubeC=cubes["C"], cube_half_size=torch.tensor([0.02] * 3), obs_mode="state",
                              agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose.create(sp_tcp)), is_grasping=lambda c_: grsp[c_.tag].clone()))
    yinfo = SPy.evaluate(fake_sp)
    G["pyr_A"], G["pyr_B"], G["pyr_C"], G["pyr_tcp"] = rA, rB, rC, sp_tcp
    for n_ in "ABC":
        G[f"pyr_static_{n_}"], G[f"pyr_grasp_{n_}"] = stat[n_], grsp[n_]
    G["pyr_success"] = yinfo["success"]
    G["pyr_extra_flat"] = common.flatten_state_dict(SPy._get_obs_extra(fake_sp, yinfo), use_torch=True)
    np.savez_compressed(OUT, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in G.items()})
    print("wrote", OUT, len(G), "arrays")


if __name__ == "__main__":
    main()
