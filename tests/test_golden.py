"""CPU: the host-side mirror against golden vectors produced by the REFERENCE's own functions
(tests/golden/make_golden.py, run in the build container where /root/reference exists; the .npz is committed)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from maniskill_b200 import utils as U
from maniskill_b200.structs import Pose

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_golden.npz"))
T = lambda k: torch.from_numpy(G[k])


def close(a, b, tol=1e-6):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    assert np.allclose(a, b, atol=tol, rtol=tol), np.abs(a - b).max()


def test_rotation_conversions():
    qa, qb, v = T("rot_qa"), T("rot_qb"), T("rot_v")
    close(U.quat_mul(qa, qb), G["rot_qmul"])
    close(U.quat_apply(qa, v), G["rot_qapply"])
    close(U.quat_to_matrix(qa), G["rot_q2m"])
    close(U.matrix_to_quat(U.quat_to_matrix(qa)), G["rot_m2q"])
    close(U.euler_xyz_to_matrix(T("rot_euler")), G["rot_euler_xyz_m"])


def test_random_quaternions_same_stream():
    torch.manual_seed(777)
    close(U.random_quaternions(16, lock_x=True, lock_y=True), G["randq_lockxy"])
    torch.manual_seed(778)
    got = U.random_quaternions(16).numpy()
    ref = G["randq_free"]
    assert np.allclose(np.minimum(np.abs(got - ref).max(1), np.abs(got + ref).max(1)), 0, atol=1e-6)


def test_clip_and_scale_action():
    close(U.clip_and_scale_action(T("cs_action"), T("cs_low"), T("cs_high")), G["cs_out"])
    close(U.clip_and_scale_action(T("cs_action")[:, :1], torch.tensor([-0.01]), torch.tensor([0.04])), G["cs_out_grip"])


def test_flatten_and_angle():
    d = dict(agent=dict(qpos=T("fl_qpos"), qvel=T("fl_qvel")), extra=dict(is_grasped=T("fl_isg"), tcp_pose=T("fl_tcp"), goal_pos=T("fl_goal")))
    out = U.flatten_state_dict(d)
    assert out.dtype == torch.float32
    close(out, G["fl_out"])
    close(U.compute_angle_between(T("ang_x1"), T("ang_x2")), G["ang_out"])


def test_matrix_to_euler_xyz():
    close(U.matrix_to_euler_xyz(U.quat_to_matrix(T("eul_q"))), G["eul_xyz_from_matrix"], 2e-6)
    # round trip with the inverse conversion
    e = T("eul_xyz_from_matrix")
    close(U.quat_to_matrix(T("eul_q")), U.euler_xyz_to_matrix(e), 2e-6)


def test_look_at():
    for name, eye, tgt in [("pick_sensor", [0.3, 0, 0.6], [-0.1, 0, 0.1]), ("pick_human", [0.6, 0.7, 0.6], [0.0, 0.0, 0.35]), ("peg_sensor", [0, -0.3, 0.2], [0, 0, 0.1])]:
        close(U.look_at(eye, tgt), G["lookat_" + name], 1e-6)


def test_pose_algebra():
    a, b = Pose(T("pose_a")), Pose(T("pose_b"))
    close((a * b).raw_pose, G["pose_mul"])
    close(a.inv().raw_pose, G["pose_inv"])
    close(a.to_transformation_matrix(), G["pose_mat"])


def _fake_pick_cube():
    from maniskill_b200.envs.pick_cube import PickCubeEnv
    qvel = T("pc_qvel")
    isg = T("pc_is_grasped")
    agent = SimpleNamespace(tcp_pose=Pose(T("pc_tcp")), robot=SimpleNamespace(get_qvel=lambda: qvel),
                            is_grasping=lambda obj: isg, is_static=lambda thr: torch.max(torch.abs(qvel[:, :-2]), 1)[0] <= thr)
    fake = SimpleNamespace(cube=SimpleNamespace(pose=Pose(T("pc_cube"))), goal_site=SimpleNamespace(pose=Pose(torch.hstack([T("pc_goal"), torch.tensor([[1.0, 0, 0, 0]]).expand(len(qvel), 4)]))),
                           agent=agent, goal_thresh=0.025, obs_mode="state")
    return PickCubeEnv, fake


def test_pick_cube_evaluate_reward_obs():
    cls, fake = _fake_pick_cube()
    info = cls.evaluate(fake)
    assert np.array_equal(info["success"].numpy(), G["pc_success"])
    assert np.array_equal(info["is_obj_placed"].numpy(), G["pc_is_obj_placed"])
    assert np.array_equal(info["is_robot_static"].numpy(), G["pc_is_robot_static"])
    close(cls.compute_dense_reward(fake, None, None, info), G["pc_reward"])
    fake.compute_dense_reward = lambda obs, action, info: cls.compute_dense_reward(fake, obs, action, info)
    close(cls.compute_normalized_dense_reward(fake, None, None, info), G["pc_reward_norm"])
    close(U.flatten_state_dict(cls._get_obs_extra(fake, info)), G["pc_extra_flat"])


def test_panda_is_grasping_and_static():
    from maniskill_b200.agents import Panda
    lf, rf = T("pg_lforce"), T("pg_rforce")
    f1, f2 = SimpleNamespace(pose=Pose(T("pg_f1"))), SimpleNamespace(pose=Pose(T("pg_f2")))
    qvel = T("pc_qvel")
    fake = SimpleNamespace(finger1_link=f1, finger2_link=f2, robot=SimpleNamespace(get_qvel=lambda: qvel),
                           scene=SimpleNamespace(get_pairwise_contact_forces=lambda a, b: lf if a is f1 else rf))
    assert np.array_equal(Panda.is_grasping(fake, None).numpy(), G["pg_is_grasping"])
    assert np.array_equal(Panda.is_static(fake, 0.2).numpy(), G["pg_is_static"])


def _fake_peg():
    from maniskill_b200.envs.peg_insertion_side import PegInsertionSideEnv as PE
    m = len(G["peg_grasped"])
    grasped = T("peg_grasped")
    hole_off = Pose(torch.hstack([T("peg_hole_off"), torch.tensor([[1.0, 0, 0, 0]]).expand(m, 4)]))
    half = T("peg_half")
    head_off = Pose(torch.hstack([half[:, :1], torch.zeros(m, 2), torch.tensor([[1.0, 0, 0, 0]]).expand(m, 4)]))
    fake = SimpleNamespace(peg=SimpleNamespace(pose=Pose(T("peg_peg"))), box=SimpleNamespace(pose=Pose(T("peg_box"))), peg_head_offsets=head_off,
                           box_hole_offsets=hole_off, box_hole_radii=T("peg_hole_radii"), peg_half_sizes=half, obs_mode="state", device=torch.device("cpu"),
                           agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose(T("peg_tcp"))), is_grasping=lambda obj, max_angle=None: grasped))
    fake.peg_head_pose = PE.peg_head_pose.fget(fake)
    fake.box_hole_pose = PE.box_hole_pose.fget(fake)
    fake.goal_pose = PE.goal_pose.fget(fake)
    fake.has_peg_inserted = lambda: PE.has_peg_inserted(fake)
    return PE, fake


def test_peg_insertion_evaluate_reward_obs():
    """mani_skill/envs/tasks/tabletop/peg_insertion_side.py:250-360 run by the reference's own code on the same synthetic states."""
    PE, fake = _fake_peg()
    close(fake.peg_head_pose.raw_pose, G["peg_head_pose"])
    close(fake.box_hole_pose.raw_pose, G["peg_box_hole_pose"])
    close(fake.goal_pose.raw_pose, G["peg_goal_pose"], 2e-6)
    info = PE.evaluate(fake)
    assert np.array_equal(info["success"].numpy(), G["peg_success"])
    assert G["peg_success"][:4].all() and not G["peg_success"].all()  # the fixture holds inserted and not-inserted pegs
    close(info["peg_head_pos_at_hole"], G["peg_head_at_hole"], 2e-6)
    close(PE.compute_dense_reward(fake, None, None, info), G["peg_reward"], 2e-5)
    close(U.flatten_state_dict(PE._get_obs_extra(fake, info)), G["peg_extra_flat"], 2e-6)


def test_open_cabinet_drawer_evaluate_reward_obs():
    """mani_skill/envs/tasks/mobile_manipulation/open_cabinet_drawer.py:221-358 run by the reference's own code."""
    from maniskill_b200.envs.open_cabinet_drawer import OpenCabinetDrawerEnv as CE
    jq = T("cab_joint_qpos")
    fake = SimpleNamespace(handle_link=SimpleNamespace(pose=Pose(T("cab_handle_pose")), angular_velocity=T("cab_ang_v"), linear_velocity=T("cab_lin_v")),
                           handle_link_pos=T("cab_handle_local"), target_qpos=T("cab_target_qpos"), obs_mode="state", device=torch.device("cpu"),
                           agent=SimpleNamespace(tcp=SimpleNamespace(pose=Pose(T("cab_tcp")))), _target_joint_qpos=lambda: jq)
    fake.handle_link_positions = lambda env_idx=None: CE.handle_link_positions(fake, env_idx)
    info = CE.evaluate(fake)
    assert np.array_equal(info["success"].numpy(), G["cab_success"])
    assert np.array_equal(info["open_enough"].numpy(), G["cab_open_enough"])
    assert G["cab_open_enough"].any() and not G["cab_open_enough"].all()
    close(info["handle_link_pos"], G["cab_handle_link_pos"], 2e-6)
    close(CE.compute_dense_reward(fake, None, None, info), G["cab_reward"], 2e-6)
    close(U.flatten_state_dict(CE._get_obs_extra(fake, info)), G["cab_extra_flat"], 2e-6)




def test_uniform_placement_sampler_same_stream_as_the_reference():
    """samplers.py:13-108 under the same global torch seed: identical points (the rejection loop consumes torch.rand identically)."""
    torch.manual_seed(31337)
    sp = U.UniformPlacementSampler([[-0.05, -0.06], [0.05, 0.06]], 16)
    pts = torch.stack([sp.sample(0.03, 100), sp.sample(0.03, 100), sp.sample(0.02, 100)])
    close(pts, G["sampler_pts"], 1e-7)
    d01 = np.linalg.norm(G["sampler_pts"][0] - G["sampler_pts"][1], axis=-1)
    assert (d01 > 0.06).all()  # the constraint the sampler enforces (and the bounds are tight enough that rejections happened)




class _FakeArticulation:
    """Just enough of structs.Articulation for the joint controllers: qpos, limits, and a capture of the drive targets."""

    def __init__(self, qpos, names):
        self.scene = SimpleNamespace(device=torch.device("cpu"))
        self.dof_names = names
        self.qpos = qpos
        self.qlimits = torch.stack([torch.full((len(names),), -3.0), torch.full((len(names),), 3.0)], 1)[None].expand(qpos.shape[0], -1, -1)
        self.sent = None

    def set_joint_drive_targets(self, targets, idx):
        self.sent = targets.clone()


def test_joint_controllers_match_the_reference_set_action():
    """pd_joint_pos.py:77-101 / 207-228 with base_controller.py:125-173: drive targets written for delta, target-delta, absolute
    and mimic control, produced by the reference's own set_action on the same inputs."""
    from maniskill_b200.agents import PDJointPosController, PDJointPosMimicController
    names = [f"j{i}" for i in range(7)]
    q = T("ctl_qpos_arm")
    art = _FakeArticulation(q, names)
    c = PDJointPosController(art, names, -0.1, 0.1, use_delta=True)
    c.reset()
    c.set_action(T("ctl_act0"))
    close(art.sent, G["ctl_delta_target"], 1e-7)
    c = PDJointPosController(art, names, -0.1, 0.1, use_delta=True, use_target=True)
    c.reset()
    c.set_action(T("ctl_act0"))
    c.set_action(T("ctl_act1"))
    close(art.sent, G["ctl_target_delta_target"], 1e-7)
    c = PDJointPosController(art, names, None, None, normalize_action=False)
    c.reset()
    c.set_action(T("ctl_act0"))
    close(art.sent, G["ctl_abs_target"], 1e-7)
    gq = T("ctl_qpos_grip")
    gart = _FakeArticulation(gq, ["f1", "f2"])
    g = PDJointPosMimicController(gart, ["f1", "f2"], -0.01, 0.04, mimic={"f2": {"joint": "f1"}})
    g.reset()
    g.set_action(T("ctl_act_grip"))
    close(gart.sent, G["ctl_mimic_target"], 1e-7)
    assert torch.equal(gart.sent[:, 0], gart.sent[:, 1])


def test_base_forward_velocity_controller_matches_the_reference():
    """pd_base_vel.py:39-73 (Fetch mobile base): forward speed and yaw rate -> velocity targets of the x / y / yaw joints."""
    from maniskill_b200.agents import PDBaseForwardVelController
    q = T("ctl_base_q")
    art = _FakeArticulation(q, ["x", "y", "yaw"])
    art.set_joint_drive_velocity_targets = lambda t, idx: setattr(art, "sent", t.clone())
    c = PDBaseForwardVelController(art, ["x", "y", "yaw"], [-1.0, -3.14], [1.0, 3.14])
    c.set_action(T("ctl_base_act"))
    close(art.sent, G["ctl_base_vel_target"], 1e-6)


def test_vector_wrapper_metrics_and_auto_reset_match_the_reference():
    """mani_skill/vector/wrappers/gymnasium.py:96-176 (`ManiSkillVectorEnv.reset` / `.step` with record_metrics and auto reset) run
    by the reference's own code on the same scripted inner environment: observations handed out, terminations / truncations, the
    running episode statistics before and after partial resets."""
    import maniskill_b200 as ms
    rew, succ, trunc = T("vec_rew"), T("vec_succ"), T("vec_trunc")
    n_t, nv = rew.shape

    class ScriptedEnv:
        max_episode_steps = None

        def __init__(self):
            self.t = 0
            self.elapsed_steps = torch.zeros(nv, dtype=torch.int32)
            self.device = torch.device("cpu")
            self.num_envs = nv

        def step(self, a):
            self.elapsed_steps = self.elapsed_steps + 1
            t = self.t
            self.t += 1
            return (torch.full((nv, 2), float(t)), rew[t].clone(), succ[t].clone(), trunc[t].clone(),
                    dict(success=succ[t].clone(), elapsed_steps=self.elapsed_steps.clone()))

        def reset(self, seed=None, options=None):
            idx = options["env_idx"] if options and "env_idx" in options else torch.arange(nv)
            self.elapsed_steps[idx] = 0
            return torch.full((nv, 2), -1.0), dict(reset=True)

    inner = ScriptedEnv()
    venv = ms.ManiSkillVectorEnv(inner, auto_reset=True, record_metrics=True)
    # the reference gets truncations from the TimeLimit wrapper inside the env; here the wrapper computes them from
    # elapsed_steps >= max_episode_steps: script the same truncation pattern through that path
    for t in range(n_t):
        venv.max_episode_steps = 1 if trunc[t].all() else 10 ** 6
        o, r, te, tr, info = venv.step(None)
        close(o, G[f"vec_obs_{t}"])
        assert np.array_equal(te.numpy(), G[f"vec_term_{t}"]) and np.array_equal(tr.numpy(), G[f"vec_truncout_{t}"])
        assert ("final_info" in info) == bool(G[f"vec_has_final_{t}"])
        ep = info["final_info"]["episode"] if "final_info" in info else info["episode"]
        assert np.array_equal(ep["success_once"].numpy(), G[f"vec_success_once_{t}"])
        close(ep["return"], G[f"vec_return_{t}"])
        assert np.array_equal(ep["episode_len"].numpy(), G[f"vec_episode_len_{t}"])
        close(ep["reward"], G[f"vec_reward_{t}"])
        close(venv.returns, G[f"vec_returns_after_{t}"])










def test_camera_parameters_of_a_mounted_camera():
    """mani_skill/utils/structs/render_camera.py:77-155 (GPU branch) on the same mount / local poses: `extrinsic_cv` [N,3,4] and
    `cam2world_gl` [N,4,4] of `CameraSensors.get_params`; intrinsics follow RenderCameraComponent.set_fovy (fy = H/2 / tan(fov/2))."""
    from maniskill_b200.render import CameraSensors, camera_desc
    m = len(G["cam_mount"])
    rows = 3
    body_view = torch.zeros(m, rows, 13)
    body_view[:, :, 3] = 1
    body_view[:, 1, :7] = T("cam_mount")
    cs = CameraSensors.__new__(CameraSensors)    # the parameter path needs no device: skip the camera-group creation
    cs.cams = [camera_desc("hand_camera", G["cam_local"][0], 128, 96, np.pi / 2, 0.01, 100, mount_row=1)]
    p = cs.get_params(body_view)["hand_camera"]
    close(p["extrinsic_cv"], G["cam_extrinsic_cv"], 2e-6)
    close(p["cam2world_gl"], G["cam_model_gl"], 2e-6)
    K = p["intrinsic_cv"]
    assert K.shape == (m, 3, 3) and np.allclose(K[0].numpy(), [[48.0, 0, 64], [0, 48.0, 48], [0, 0, 1]], atol=1e-4)
    # a mounted camera is recomputed every call, a fixed one is cached
    body_view[:, 1, 0] += 1.0
    assert not torch.allclose(cs.get_params(body_view)["hand_camera"]["extrinsic_cv"], p["extrinsic_cv"])


def test_end_effector_controllers_and_the_ik_step_match_the_reference():
    """mani_skill/agents/controllers/pd_ee_pose.py:85-99,229-263 (action clipping: rotation clipped by norm and scaled by `rot_lower`;
    target pose in the frame root_translation[:root_aligned_body_rotation]) and utils/kinematics.py:197-260 (GPU branch: delta pose
    from a target pose, damped least squares / pseudo-inverse on a given Jacobian), run by the reference's own code on the same inputs."""
    from maniskill_b200.agents import PDEEPosController, PDEEPoseController
    from maniskill_b200.kinematics import Kinematics
    act, prev = T("ee_act"), Pose(T("ee_prev"))
    fake = SimpleNamespace(normalize_action=True, action_low=torch.full((6,), -0.1), action_high=torch.full((6,), 0.1), rot_lower=-0.1, use_delta=True)
    scaled = PDEEPoseController._preprocess_action(fake, act)
    close(scaled, G["ee_scaled"], 1e-6)
    assert (np.linalg.norm(G["ee_scaled"][:, 3:], axis=1) <= 0.1 + 1e-6).all() and (np.linalg.norm(G["ee_act"][:3, 3:], axis=1) > 1).all()
    close(PDEEPoseController.compute_target_pose(fake, prev, scaled).raw_pose, G["ee_target_pose"], 2e-6)
    close(PDEEPosController.compute_target_pose(fake, prev, scaled[:, :3]).raw_pose, G["ee_target_pos_only"], 1e-6)
    J, q0 = T("ee_J"), T("ee_q0")
    kin = Kinematics.__new__(Kinematics)
    kin.chain = SimpleNamespace(forward=lambda q: (None, None, J))
    kin.chain_dof_idx, kin.qmask, kin.device = torch.arange(7), torch.ones(7, dtype=torch.bool), torch.device("cpu")
    delta_t = PDEEPosController._delta_from_target(fake, Pose(T("ee_target_pose")), Pose(T("ee_cur")))
    for name, cfg in (("lm", dict(type="levenberg_marquardt", alpha=1.0)), ("pinv", dict(type="pseudo_inverse", alpha=0.5))):
        close(kin.compute_ik(delta_t, q0, cfg), G[f"ee_ik_target_{name}"], 2e-4)
        close(kin.compute_ik(scaled, q0, cfg), G[f"ee_ik_delta_{name}"], 2e-4)
    with pytest.raises(NotImplementedError):
        kin.compute_ik(scaled, q0, dict(type="newton"))


def test_record_episode_writes_what_the_reference_recorder_writes(tmp_path):
    """mani_skill/utils/wrappers/record.py:356-756 run by the reference's own code (h5py replaced by a dict-backed fake) on a scripted
    3-sub-scene env through full and partial resets: every dataset of every flushed episode -- order, keys, shapes, dtypes, values --
    and the per-episode JSON fields."""
    from maniskill_b200.trajectory import RecordEpisode, load_trajectories
    obs_s, rew_s, succ_s, state_s, act_s = T("rec_obs"), T("rec_rew"), T("rec_succ"), T("rec_state"), T("rec_act")
    nr = obs_s.shape[1]

    class Scripted:
        obs_mode, _reward_mode, control_mode, max_episode_steps, action_dim = "state", "dense", "pd_joint_delta_pos", None, 2

        def __init__(self):
            self.k, self.t, self.num_envs = 0, 0, nr
            self._episode_seed = np.array([5, 6, 7])
            self.cur = torch.zeros(nr, 5)

        def get_state_dict(self):
            return dict(actors=dict(cube=self.cur.clone()), articulations=dict(panda=self.cur.clone() * 2))

        def reset(self, seed=None, options=None):
            idx = torch.arange(nr) if not options or "env_idx" not in options else torch.as_tensor(options["env_idx"])
            self.cur[idx] = state_s[self.k][idx]
            self.k += 1
            return obs_s[self.k - 1].clone(), dict(reconfigure=False)

        def step(self, a):
            self.cur = state_s[self.k].clone()
            self.k += 1
            self.t += 1
            t = self.t - 1
            return obs_s[self.k - 1].clone(), rew_s[t].clone(), succ_s[t].clone(), torch.zeros(nr, dtype=torch.bool), dict(success=succ_s[t].clone())

    rec = RecordEpisode(Scripted(), str(tmp_path), env_id="Scripted-v0")
    rec.reset(seed=5)
    for t in range(3):
        rec.step(act_s[t])
    rec.reset(options=dict(env_idx=torch.tensor([1])))
    for t in range(3, 5):
        rec.step(act_s[t])
    rec.reset(options=dict(env_idx=torch.tensor([0, 2])))
    for t in range(5, 7):
        rec.step(act_s[t])
    rec.reset()
    rec.step(act_s[7])
    rec.flush_trajectory()
    rec._dump()
    meta, trajs = load_trajectories(str(tmp_path / "trajectory"))
    flat = {}
    from maniskill_b200.trajectory import _flatten
    for name, tr in trajs.items():
        _flatten(name, tr, flat)
    ref = {k[len("recout/"):]: G[k] for k in G.files if k.startswith("recout/")}
    assert set(flat) == set(ref) and len(ref) == 72
    for k in ref:
        assert flat[k].dtype == ref[k].dtype and flat[k].shape == ref[k].shape, k
        assert np.array_equal(flat[k], ref[k]), k
    eps = meta["episodes"]
    assert [e["elapsed_steps"] for e in eps] == G["rec_episode_steps"].tolist()
    assert [e["episode_seed"] for e in eps] == G["rec_episode_seed"].tolist()
    assert [e["success"] for e in eps] == G["rec_episode_success"].tolist()
    assert [e["episode_id"] for e in eps] == list(range(9))


def test_sensor_data_to_pointcloud_matches_the_reference():
    """mani_skill/envs/utils/observations/observations.py:16-68 on the same two-camera synthetic targets: world-frame homogeneous
    points (w = 0 on background), colours and ids, cameras concatenated along the point axis, `sensor_data` emptied."""
    from maniskill_b200.observations import sensor_data_to_pointcloud
    obs = dict(sensor_data={}, sensor_param={})
    for ci, uid in enumerate(("base_camera", "hand_camera")):
        obs["sensor_data"][uid] = dict(rgb=T(f"pcd_in_{ci}_rgb"), position=T(f"pcd_in_{ci}_position"), segmentation=T(f"pcd_in_{ci}_segmentation"))
        obs["sensor_param"][uid] = dict(cam2world_gl=T(f"pcd_in_{ci}_cam2world"))
    pos_before = obs["sensor_data"]["base_camera"]["position"].clone()
    out = sensor_data_to_pointcloud(obs)
    assert out["sensor_data"] == {} and set(out["pointcloud"]) == {"xyzw", "rgb", "segmentation"}
    close(out["pointcloud"]["xyzw"], G["pcd_out_xyzw"], 1e-6)
    for k in ("rgb", "segmentation"):
        assert out["pointcloud"][k].dtype == torch.from_numpy(G[f"pcd_out_{k}"]).dtype and np.array_equal(out["pointcloud"][k].numpy(), G[f"pcd_out_{k}"])
    assert torch.equal(T("pcd_in_0_position"), pos_before)       # the render target is left in millimetres


def test_flatten_rgbd_observation_wrapper_matches_the_reference():
    """mani_skill/utils/wrappers/flatten.py:42-77 on the same two-camera observation in three settings (separate depth, merged
    "rgbd" -- uint8 colour promoted to the int16 of depth --, colour only without state)."""
    from maniskill_b200.wrappers import FlattenRGBDObservationWrapper as W

    def obs():
        st = T("flat_in_state_parts")
        cams = {uid: dict(rgb=T(f"flat_in_rgb_{i}"), depth=T(f"flat_in_depth_{i}")) for i, uid in enumerate(("base_camera", "hand_camera"))}
        return dict(agent=dict(qpos=st[:, :9], qvel=st[:, 9:18]), extra=dict(is_grasped=st[:, 18] > 0.5, tcp_pose=st[:, 19:]),
                    sensor_param=dict(base_camera={}, hand_camera={}), sensor_data=cams)

    for tag, kw in (("sep", dict(include_rgb=True, include_depth=True, sep_depth=True, include_state=True)),
                    ("merged", dict(include_rgb=True, include_depth=True, sep_depth=False, include_state=True)),
                    ("rgbonly", dict(include_rgb=True, include_depth=False, sep_depth=True, include_state=False))):
        src = obs()
        out = W.observation(SimpleNamespace(**kw), src)
        assert sorted(out.keys()) == G[f"flat_keys_{tag}"].tolist()
        for k, v in out.items():
            ref = G[f"flat_{tag}_{k}"]
            assert v.numpy().dtype == ref.dtype and np.array_equal(v.numpy(), ref), (tag, k)
        assert "sensor_data" in src       # the caller's dict is left intact


def test_reset_seed_derivation_and_per_sub_scene_streams_match_the_reference():
    """mani_skill/envs/sapien_env.py:980-1016 (`_set_main_rng`, `_set_episode_rng`) + envs/utils/randomization/batched_rng.py run by the
    reference's own code: reset(seed=7) fans one seed out to every sub-scene, an unseeded reset keeps the streams running, enhanced
    determinism re-seeds exactly the sub-scenes being reset from their main generators, a list of seeds is taken as is -- compared
    through the seeds and through draws from the resulting generators."""
    from maniskill_b200.envs.base_env import BaseEnv
    nv = 4
    fe = SimpleNamespace(num_envs=nv, _main_seed=None, _enhanced_determinism=False, _episode_seed=np.zeros(nv, dtype=np.int64), _batched_episode_rng=None,
                         _batched_main_rng=None, _episode_rng=None)

    def check(tag):
        assert np.array_equal(np.asarray(fe._main_seed, dtype=np.int64), G[f"rng_{tag}_main_seed"]), tag
        assert np.array_equal(np.asarray(fe._episode_seed, dtype=np.int64), G[f"rng_{tag}_episode_seed"]), tag
        assert np.array_equal(fe._batched_episode_rng.uniform(0, 1, size=(2,)), G[f"rng_{tag}_draw"]), tag
        assert np.array_equal(fe._episode_rng.normal(0, 0.02, (2, 3)), G[f"rng_{tag}_env0_normal"]), tag

    BaseEnv._set_main_rng(fe, 7)
    BaseEnv._set_episode_rng(fe, 7, torch.arange(nv))
    check("a")
    BaseEnv._set_main_rng(fe, None)
    BaseEnv._set_episode_rng(fe, None, torch.arange(nv))
    check("b")
    fe._enhanced_determinism = True
    BaseEnv._set_main_rng(fe, None)
    BaseEnv._set_episode_rng(fe, None, torch.tensor([1, 3]))
    check("c")
    BaseEnv._set_main_rng(fe, [11, 12, 13, 14])
    BaseEnv._set_episode_rng(fe, [11, 12, 13, 14], torch.arange(nv))
    check("d")
    assert len(set(G["rng_a_episode_seed"].tolist())) == nv and not np.array_equal(G["rng_c_episode_seed"], G["rng_b_episode_seed"])


def test_step_termination_and_reward_dispatch_match_the_reference():
    """mani_skill/envs/sapien_env.py:648-700,1042-1071 run by the reference's own code on scripted task hooks, for every combination of
    success / fail being reported and every reward mode: reward, terminated (success | fail), truncated (never set by BaseEnv), and the
    step counter.  (Fail-only + sparse is left out: the reference negates a bool tensor there and raises.)"""
    from maniskill_b200.envs.base_env import BaseEnv
    succ, fail, dense = T("step_succ"), T("step_fail"), T("step_dense")
    ns = len(succ)
    for tag, keys in (("sf", ("success", "fail")), ("s", ("success",)), ("f", ("fail",)), ("none", ())):
        for mode in ("sparse", "dense", "normalized_dense", "none"):
            if (tag, mode) == ("f", "sparse"):
                continue
            info0 = {k: dict(success=succ, fail=fail)[k].clone() for k in keys}
            fs = SimpleNamespace(num_envs=ns, device=torch.device("cpu"), _elapsed_steps=torch.zeros(ns, dtype=torch.int32), _reward_mode=mode, _fused=None,
                                 _state_version=0, _step_action=lambda a: a, get_info=lambda: dict(info0), _obs_mode="none", _epilogue_runner=None,
                                 compute_dense_reward=lambda obs, action, info: dense * 5, compute_normalized_dense_reward=lambda obs, action, info: dense)
            fs.get_reward = lambda obs, action, info: BaseEnv.get_reward(fs, obs, action, info)
            fs.compute_sparse_reward = lambda obs, action, info: BaseEnv.compute_sparse_reward(fs, obs, action, info)
            fs._epilogue = lambda action: BaseEnv._epilogue(fs, action)
            fs._obs_from_core = lambda core: BaseEnv._obs_from_core(fs, core)
            o, r, te, tr, i = BaseEnv.step(fs, torch.zeros(ns, 3))
            close(r.float(), G[f"step_{tag}_{mode}_reward"], 1e-7)
            assert np.array_equal(te.numpy(), G[f"step_{tag}_{mode}_terminated"]) and np.array_equal(tr.numpy(), G[f"step_{tag}_{mode}_truncated"])
            assert int(fs._elapsed_steps[0]) == 1
    fs = SimpleNamespace(num_envs=ns, device=torch.device("cpu"))
    assert torch.equal(BaseEnv.compute_sparse_reward(fs, None, None, dict(fail=fail)), -fail.float())      # the documented intent of the raising line


def test_tile_images_and_camera_images_match_the_reference():
    """mani_skill/utils/visualization/misc.py:54-115 (`tile_images`: batched and single images, mixed sizes, two rows) and
    mani_skill/sensors/camera.py:256-294 (`camera_observations_to_images`: depth / position normalised to grey, hashed segmentation
    colours) run by the reference's own code on the same inputs."""
    from maniskill_b200.visualization import camera_observations_to_images, tile_images
    for tag, n in (("a", 3), ("b", 3), ("c", 4)):
        out = tile_images([T(f"tile_{tag}_in{i}") for i in range(n)])
        assert out.numpy().dtype == G[f"tile_{tag}_out"].dtype and np.array_equal(out.numpy(), G[f"tile_{tag}_out"]), tag
    two = tile_images([T(f"tile_b_in{i}") for i in range(3)] + [T("tile_b_in0")], nrows=2)
    assert np.array_equal(two.numpy(), G["tile_b_out_2rows"])
    obs = {k: T(f"camimg_in_{k}") for k in ("rgb", "depth", "segmentation", "position")}
    out = camera_observations_to_images(obs)
    assert set(out) == {"rgb", "depth", "segmentation", "position"}
    for k, v in out.items():
        assert v.numpy().dtype == G[f"camimg_out_{k}"].dtype and np.array_equal(v.numpy(), G[f"camimg_out_{k}"]), k


def test_velocity_controllers_match_the_reference_set_action():
    """pd_joint_vel.py:38-42 and pd_joint_pos_vel.py:43-67 with base_controller.py:125-173, produced by the reference's own set_action:
    clipped / scaled velocity targets, and for the position + velocity controller (target-delta positions over two actions) both halves."""
    from maniskill_b200.agents import PDJointPosVelController, PDJointVelController
    names = [f"j{i}" for i in range(7)]

    class Art(_FakeArticulation):
        def __init__(self, qpos, names):
            super().__init__(qpos, names)
            self.vel_sent = None

        def set_joint_drive_velocity_targets(self, targets, idx):
            self.vel_sent = targets.clone()

    art = Art(T("ctl_qpos_arm"), names)
    v = PDJointVelController(art, names, -1.0, 1.0)
    v.set_action(T("ctl_vel_act"))
    close(art.vel_sent, G["ctl_vel_target"], 1e-7)
    assert (np.abs(G["ctl_vel_target"]) <= 1 + 1e-6).all() and (np.abs(G["ctl_vel_act"]) > 1).any()
    art.scene.world = SimpleNamespace(target_qvel=torch.zeros(len(art.qpos), 7))       # what reset() clears
    art.scene.BUF_TARGET_QVEL, art.scene._dirty, art._rows = 64, 0, torch.arange(len(art.qpos))
    pv = PDJointPosVelController(art, names, -0.1, 0.1, use_delta=True, use_target=True)
    pv.reset()
    assert pv.action_dim == 14
    pv.set_action(T("ctl_pv_act0"))
    pv.set_action(T("ctl_pv_act1"))
    close(art.sent, G["ctl_pv_pos_target"], 1e-7)
    close(art.vel_sent, G["ctl_pv_vel_target"], 1e-7)




