"""CPU (-m "not gpu"): host logic of the BaseEnv mirror on the emulated backend (tests/emu_world.py).
Restates the semantic checks of the reference's own tests:
  tests/test_envs.py:151-184 (same-seed determinism), tests/test_gpu_envs.py:245-270 (partial reset isolation),
  tests/test_gpu_envs.py:272-285 (truncation after max_episode_steps), tests/test_sim_state.py:10-32 (state vector
  width 70 for PickCube-v1 and set/get round trip), tests/test_envs.py:56-76 (obs shapes/dtypes)."""
import numpy as np
import pytest
import torch

import maniskill_b200 as ms
from emu_world import EmuBackendWorld


def make(n=4, **kw):
    return ms.make("PickCube-v1", num_envs=n, obs_mode="state", world_factory=EmuBackendWorld, **kw)


def test_obs_shape_and_action_space():
    env = make(3)
    obs, info = env.reset(seed=0)
    assert obs.shape == (3, 42) and obs.dtype == torch.float32
    assert env.action_dim == 8
    assert np.allclose(env.single_action_space_low, -1) and np.allclose(env.single_action_space_high, 1)
    o, r, te, tr, info = env.step(torch.zeros(3, 8))
    assert o.shape == (3, 42) and r.shape == (3,) and te.dtype == torch.bool and tr.dtype == torch.bool
    for k in ("success", "is_obj_placed", "is_robot_static", "is_grasped", "elapsed_steps"):
        assert k in info


def test_same_seed_is_deterministic():
    env = make(2)
    a = [2 * torch.rand(2, 8, generator=torch.Generator().manual_seed(i)) - 1 for i in range(6)]
    obs1, _ = env.reset(seed=2000)
    tr1 = [env.step(x)[0].clone() for x in a]
    obs2, _ = env.reset(seed=2000)
    tr2 = [env.step(x)[0].clone() for x in a]
    assert torch.allclose(obs1, obs2, atol=1e-4)
    for x, y in zip(tr1, tr2):
        assert torch.allclose(x, y, atol=1e-4)
    obs3, _ = env.reset(seed=2001)
    assert not torch.allclose(obs1, obs3, atol=1e-4)


def test_partial_reset_only_touches_selected_envs():
    env = make(4)
    env.reset(seed=0)
    for _ in range(3):
        obs, *_ = env.step(2 * torch.rand(4, 8) - 1)
    before = obs.clone()
    obs2, _ = env.reset(options=dict(env_idx=torch.tensor([1, 3])))
    assert torch.allclose(obs2[[0, 2]], before[[0, 2]], atol=1e-4)
    assert not torch.allclose(obs2[[1, 3]], before[[1, 3]], atol=1e-4)
    assert env.elapsed_steps.tolist() == [3, 0, 3, 0]


def test_truncation_and_auto_reset():
    env = make(2)
    venv = ms.ManiSkillVectorEnv(env)
    venv.reset(seed=0)
    for i in range(50):
        obs, r, te, tr, info = venv.step(torch.zeros(2, 8))
        if i < 49:
            assert not tr.any()
    assert tr.all()
    assert "final_info" in info and env.elapsed_steps.tolist() == [0, 0]


def test_state_vector_width_and_round_trip():
    env = make(2)
    env.reset(seed=1)
    state = env.get_state_dict()
    flat = env.get_state()
    # 3 actors x 13 + (13 + 2*9)  (tests/test_sim_state.py:20-32)
    assert flat.shape == (2, 13 * 3 + 13 + 9 * 2)
    saved = {k: {n: v.clone() for n, v in d.items()} for k, d in state.items()}
    obs0 = env.get_obs()
    for _ in range(3):
        env.step(2 * torch.rand(2, 8) - 1)
    env.set_state_dict(saved)
    assert torch.allclose(env.get_obs()[:, :18], obs0[:, :18], atol=1e-4)  # qpos, qvel
    assert torch.allclose(env.get_state(), torch.hstack([torch.hstack(list(saved["actors"].values())), torch.hstack(list(saved["articulations"].values()))]), atol=1e-5)


def test_reset_places_cube_and_goal_in_range():
    env = make(8)
    env.reset(seed=3)
    cube = env.cube.pose.p
    goal = env.goal_site.pose.p
    assert (cube[:, :2].abs() <= 0.1 + 1e-6).all() and torch.allclose(cube[:, 2], torch.full((8,), 0.02), atol=1e-6)
    assert (goal[:, :2].abs() <= 0.1 + 1e-6).all() and (goal[:, 2] >= 0.02 - 1e-6).all() and (goal[:, 2] <= 0.32 + 1e-6).all()
    q = env.agent.robot.qpos
    assert torch.allclose(q[:, 7:], torch.full((8, 2), 0.04))


def test_unsupported_modes_raise():
    with pytest.raises(NotImplementedError):
        ms.make("PickCube-v1", num_envs=1, obs_mode="rgb+albedo", world_factory=EmuBackendWorld)   # texture the minimal shader does not write
    with pytest.raises(NotImplementedError):
        ms.make("PickCube-v1", num_envs=1, control_mode="pd_ee_twist", world_factory=EmuBackendWorld)  # no such mode for the Panda


def test_peg_insertion_side_heterogeneous_envs():
    """tests/test_sim_state.py:53-65 analogue + per-env geometry (peg_insertion_side.py:114-120)."""
    env = ms.make("PegInsertionSide-v1", num_envs=3, obs_mode="state", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    assert obs.shape == (3, 9 + 9 + 7 + 7 + 3 + 7 + 1)
    hs = env.peg_half_sizes
    assert (hs[:, 0] >= 0.085).all() and (hs[:, 0] <= 0.125).all() and (hs[:, 1] >= 0.015).all() and (hs[:, 1] <= 0.025).all()
    assert len(torch.unique(hs[:, 0])) == 3  # every env has its own peg
    # same draws as the reference: RandomState(2022 + i).uniform(0.085, 0.125) is the first draw of env i
    assert np.allclose(hs[:, 0].numpy(), [np.random.RandomState(2022 + i).uniform(0.085, 0.125) for i in range(3)], atol=1e-7)
    for _ in range(20):
        obs, r, te, tr, info = env.step(torch.zeros(3, 8))
    # pegs lie flat on the table: z = radius (within the soft-contact sag)
    assert torch.allclose(env.peg.pose.p[:, 2], hs[:, 1], atol=1.5e-3)
    assert not info["success"].any() and torch.isfinite(r).all()
    # goal geometry helper: the hole frame sits at the box centre plus the per-env offset
    assert torch.allclose((env.box.pose.inv() * env.box_hole_pose).p[:, 0], torch.zeros(3), atol=1e-6)


def test_open_cabinet_drawer_standin():
    env = ms.make("OpenCabinetDrawer-v1", num_envs=2, obs_mode="state", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    assert env.action_dim == 13 and obs.shape == (2, 15 + 15 + 7 + 3 + 1 + 3)
    q = env.agent.robot.qpos
    names = env.agent.robot.dof_names
    d = torch.sqrt(q[:, names.index("root_x_axis_joint")] ** 2 + q[:, names.index("root_y_axis_joint")] ** 2)
    assert ((d >= 1.6 - 1e-4) & (d <= 1.8 + 1e-4)).all()  # robot 1.6-1.8 m from the cabinet (open_cabinet_drawer.py:262-267)
    r0 = env.step(torch.zeros(2, 13))[1]
    # pull the target drawer open with a generalised force: reward goes up, open_enough / success flip
    qf = torch.zeros(2, env.cabinet.dof)
    qf[torch.arange(2), env._target_dof] = 40.0
    env.cabinet.set_qf(qf)
    env.scene._gpu_apply_all()
    for _ in range(40):
        obs, r, te, tr, info = env.step(torch.zeros(2, 13))
    assert (env._target_joint_qpos() > 0.25).all() and info["open_enough"].all()
    assert (r > r0).all()
    env.cabinet.set_qf(torch.zeros(2, env.cabinet.dof))
    env.scene._gpu_apply_all()
    for _ in range(10):
        obs, r, te, tr, info = env.step(torch.zeros(2, 13))
    assert info["success"].all() and torch.allclose(r, torch.ones(2))
    # the goal marker follows the handle of the chosen drawer (:294-305)
    assert torch.allclose(env.handle_link_goal.pose.p, info["handle_link_pos"], atol=1e-5)


@pytest.mark.parametrize("mode", ["state_dict", "rgbd-like"])
def test_pick_cube_obs_from_fused_vector_matches_obs_dict(mode):
    """`_obs_from_fused` (used when the fused control step serves a visual observation mode) rebuilds exactly the agent / extra
    entries `get_obs` produces, from the flattened state vector."""
    from maniskill_b200 import utils as U
    env = ms.make("PickCube-v1", num_envs=5, obs_mode="state_dict", device="cpu", world_factory=EmuBackendWorld)
    env.reset(seed=2)
    g = torch.Generator().manual_seed(0)
    for _ in range(3):
        env.step(2 * torch.rand((5, env.action_dim), generator=g) - 1)
    info = env.get_info()
    ref = env._get_obs_state_dict(info)
    vec = U.flatten_state_dict(ref)
    if mode != "state_dict":
        env._obs_mode = "rgbd"  # extra drops the privileged entries (pick_cube.py:132-145)
        ref = dict(agent=env._get_obs_agent(), extra=env._get_obs_extra(info))
        assert "obj_pose" not in ref["extra"]
    got = env._obs_from_fused(vec, info)
    assert set(got.keys()) == set(ref.keys())
    for grp in ref:
        assert set(got[grp].keys()) == set(ref[grp].keys()), grp
        for k in ref[grp]:
            assert got[grp][k].dtype == ref[grp][k].dtype and torch.equal(got[grp][k], ref[grp][k]), (grp, k)


def test_net_contact_force_on_resting_cube_is_its_weight():
    """Body-net impulse query (base.py:116-136): a cube at rest on the table feels m g from the table and nothing else; the net
    query equals the sum of the pairwise queries against everything the cube touches."""
    env = ms.make("PickCube-v1", num_envs=3, obs_mode="state", device="cpu", world_factory=EmuBackendWorld)
    env.reset(seed=0)
    for _ in range(10):
        env.step(torch.zeros(3, env.action_dim))
    f = env.cube.get_net_contact_forces()
    mass = 1000 * 0.04**3  # density x volume of the 4 cm cube (pick_cube.py:77-84)
    assert f.shape == (3, 3)
    assert torch.allclose(f[:, 2], torch.full((3,), mass * 9.81), rtol=2e-2)
    assert f[:, :2].abs().max() < 1e-2
    table = env.scene.actors["table-workspace"]
    assert torch.allclose(f, env.scene.get_pairwise_contact_forces(env.cube, table), atol=1e-6)






def test_scripted_pick_and_lift_with_the_ee_controller():
    """End to end through the public env API on the emulated device code: approach from above, close the gripper, lift 15 cm.  The
    cube comes along (patch friction of the finger pads, mimic tendon, soft contacts), `is_grasped` turns on and the staged reward
    grows (pick_cube.py:161-191)."""
    n = 2
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode="pd_ee_target_delta_pos", world_factory=EmuBackendWorld)
    env.reset(seed=4)
    ctrl = env.agent.controller.controllers["arm"]
    base = torch.tensor([-0.615, 0.0, 0.0])
    cube0 = env.cube.pose.p.clone()

    def go_to(target, steps, grip, max_step=0.03):
        for _ in range(steps):
            a = torch.zeros(n, 4)
            a[:, :3] = (target - ctrl._target_pose.p).clamp(-max_step, max_step) / 0.1
            a[:, 3] = grip
            out = env.step(a)
        return out

    # the cube's yaw is random: a 4 cm cube fits between the open fingers (8 cm) at any yaw
    go_to(cube0 - base + torch.tensor([0.0, 0.0, 0.10]), 12, 1.0)
    go_to(cube0 - base, 12, 1.0)
    obs, r_closed, *_ = go_to(cube0 - base, 8, -1.0)
    assert env.evaluate()["is_grasped"].all()
    obs, r_lift, te, tr, info = go_to(cube0 - base + torch.tensor([0.0, 0.0, 0.15]), 20, -1.0, max_step=0.015)
    lifted = env.cube.pose.p[:, 2] - cube0[:, 2]
    assert (lifted > 0.12).all(), lifted
    assert info["is_grasped"].all()
    assert (torch.linalg.norm(env.cube.pose.p - env.agent.tcp.pose.p, dim=1) < 0.015).all()   # still between the fingers
    assert (r_closed > 0.2).all()  # reaching + grasp bonus of the normalised dense reward


def test_modes_and_configs_the_reference_accepts():
    """obs_mode state_dict / none, reward_mode sparse / none / dense, a different sim / control frequency, num_envs = 1 with an
    unbatched action, explicit per-env seeds (sapien_env.py:214-245, 857-930, 1042-1071)."""
    env = ms.make("PickCube-v1", num_envs=2, obs_mode="state_dict", reward_mode="sparse", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=[3, 4])
    assert set(obs.keys()) == {"agent", "extra"} and obs["agent"]["qpos"].shape == (2, 9) and obs["extra"]["is_grasped"].dtype == torch.bool
    o, r, te, tr, info = env.step(torch.zeros(2, 8))
    assert torch.equal(r, info["success"].float())          # sparse reward = success (sapien_env.py:1079-1082)
    env_b = ms.make("PickCube-v1", num_envs=2, obs_mode="state", world_factory=EmuBackendWorld)
    ob_b, _ = env_b.reset(seed=[3, 4])
    assert torch.allclose(ob_b[:, :9], obs["agent"]["qpos"], atol=1e-6)   # the same seeds give the same episode whatever the obs mode
    env_n = ms.make("PickCube-v1", num_envs=1, obs_mode="none", reward_mode="none", world_factory=EmuBackendWorld)
    obs_n, _ = env_n.reset(seed=0)
    o, r, te, tr, info = env_n.step(torch.zeros(8))           # unbatched action for a single sub-scene
    assert o == {} and float(r.abs().sum()) == 0.0 and te.shape == (1,)
    env_d = ms.make("PickCube-v1", num_envs=2, obs_mode="state", reward_mode="dense", sim_config=dict(sim_freq=200, control_freq=50), world_factory=EmuBackendWorld)
    env_d.reset(seed=1)
    assert env_d._sim_steps_per_control == 4
    o, r, te, tr, info = env_d.step(torch.zeros(2, 8))
    env_nd = ms.make("PickCube-v1", num_envs=2, obs_mode="state", sim_config=dict(sim_freq=200, control_freq=50), world_factory=EmuBackendWorld)
    env_nd.reset(seed=1)
    o2, r2, *_ = env_nd.step(torch.zeros(2, 8))
    assert torch.allclose(r, 5 * r2, atol=1e-5)               # normalized_dense = dense / 5 (pick_cube.py:188-191)
    with pytest.raises(ValueError):
        ms.make("PickCube-v1", num_envs=1, sim_config=dict(sim_freq=100, control_freq=30), world_factory=EmuBackendWorld)


def test_reset_to_env_states_restores_a_saved_episode():
    """reset(options={"reset_to_env_states": ...}) (sapien_env.py:905-913): the saved state comes back, for the selected sub-scenes only."""
    env = make(3)
    env.reset(seed=9)
    for _ in range(4):
        env.step(2 * torch.rand(3, 8, generator=torch.Generator().manual_seed(1)) - 1)
    saved = env.get_state_dict()
    saved = {k: {n: v.clone() for n, v in d.items()} for k, d in saved.items()}
    flat_saved = env.get_state().clone()
    for _ in range(5):
        env.step(2 * torch.rand(3, 8) - 1)
    moved = env.get_state().clone()
    assert not torch.allclose(moved, flat_saved, atol=1e-4)
    env.reset(options=dict(env_idx=torch.tensor([0, 2]), reset_to_env_states=dict(env_states={k: {n: v[[0, 2]] for n, v in d.items()} for k, d in saved.items()})))
    now = env.get_state()
    assert torch.allclose(now[[0, 2]], flat_saved[[0, 2]], atol=1e-5)
    assert torch.allclose(now[1], moved[1], atol=1e-5)
    assert env.elapsed_steps.tolist()[0] == 0 and env.elapsed_steps.tolist()[1] == 9










def test_seeded_sequence_reset_with_enhanced_determinism():
    """tests/test_envs.py:166-184 of the reference: with `enhanced_determinism=True` the unseeded resets inside a rollout draw their
    episode seeds from the main generator, so the whole sequence (17 steps through episodes of 5) repeats after `reset(seed=2000)`;
    without it the unseeded resets continue the running streams, and the layouts after the same number of resets differ between the
    two settings."""
    N = 17
    g = torch.Generator().manual_seed(0)
    actions = [2 * torch.rand(1, 8, generator=g) - 1 for _ in range(N)]

    def rollout(env):
        obs, _ = env.reset(seed=2000)
        for a in actions:
            obs, _, _, _, _ = env.step(a)
            if int(env.elapsed_steps[0]) >= 5:       # TimeLimit of the reference's gym.make(max_episode_steps=5)
                obs, _ = env.reset()
        return obs.clone()

    env = ms.make("PickCube-v1", num_envs=1, obs_mode="state", world_factory=EmuBackendWorld, enhanced_determinism=True)
    first, again = rollout(env), rollout(env)
    assert torch.allclose(first, again, atol=1e-5)
    seeds_det = env._episode_seed.copy()
    plain = ms.make("PickCube-v1", num_envs=1, obs_mode="state", world_factory=EmuBackendWorld)
    p1 = rollout(plain)
    assert plain._episode_seed[0] == 2000 and seeds_det[0] != 2000          # episode seed re-drawn only under enhanced determinism
    assert not torch.allclose(p1, first, atol=1e-4)


def test_reconfigure_rebuilds_the_scene():
    """sapien_env.py:895-916 + tests/test_gpu_envs.py:145-155: `reset(options=dict(reconfigure=True))` rebuilds the scene -- tasks that draw
    geometry while building get new draws (PegInsertionSide: per-sub-scene peg sizes), seeded reconfiguration is reproducible, a partial
    reset cannot reconfigure, and `reconfiguration_freq=1` rebuilds on every reset."""
    env = ms.make("PegInsertionSide-v1", num_envs=3, obs_mode="state", world_factory=EmuBackendWorld)
    env.reset(seed=0)
    sizes0 = env.peg_half_sizes.clone()
    world0 = env.scene.world
    obs, info = env.reset(seed=5, options=dict(reconfigure=True))
    assert info["reconfigure"] and env.scene.world is not world0
    sizes1 = env.peg_half_sizes.clone()
    assert not torch.allclose(sizes0, sizes1)
    for _ in range(3):
        obs, r, te, tr, info = env.step(torch.zeros(3, 8))
    assert torch.isfinite(obs).all() and int(env.elapsed_steps[0]) == 3
    obs2, info2 = env.reset(seed=5, options=dict(reconfigure=True))
    assert torch.allclose(env.peg_half_sizes, sizes1) and torch.allclose(obs2, env.reset(seed=5, options=dict(reconfigure=True))[0], atol=1e-5)
    _, info3 = env.reset(seed=5)
    assert not info3["reconfigure"] and torch.allclose(env.peg_half_sizes, sizes1)      # a plain reset keeps the geometry
    with pytest.raises(RuntimeError, match="partial reset and reconfigure"):
        env.reset(options=dict(reconfigure=True, env_idx=torch.tensor([0])))
    every = ms.make("PegInsertionSide-v1", num_envs=2, obs_mode="state", world_factory=EmuBackendWorld, reconfiguration_freq=1)
    a = every.peg_half_sizes.clone()
    _, info = every.reset()
    assert info["reconfigure"] and not torch.allclose(every.peg_half_sizes, a)
    b = every.peg_half_sizes.clone()
    _, info = every.reset()
    assert info["reconfigure"] and not torch.allclose(every.peg_half_sizes, b)


def test_tabletop_tasks_accept_the_wrist_camera_panda():
    """SUPPORTED_ROBOTS of the tabletop tasks (pick_cube.py:40 ...): `robot_uids="panda_wristcam"` loads the v3 Panda (camera link on the
    hand), rests with the last arm joint turned the other way (table/scene_builder.py:104-108) and adds the hand camera to the sensors;
    unknown robots are refused."""
    env = ms.make("PickCube-v1", num_envs=2, obs_mode="state", robot_uids="panda_wristcam", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    assert obs.shape == (2, 42) and "panda_wristcam" in env.scene.articulations and "camera_link" in env.agent.robot.links_map
    q = env.agent.robot.get_qpos()
    assert (q[:, 6] + np.pi / 4).abs().max() < 0.1
    ref = ms.make("PickCube-v1", num_envs=2, obs_mode="state", world_factory=EmuBackendWorld)
    ref.reset(seed=0)
    assert (ref.agent.robot.get_qpos()[:, 6] - np.pi / 4).abs().max() < 0.1
    assert torch.allclose(ref.cube.pose.p, env.cube.pose.p)                      # the same layout draws
    for _ in range(3):
        obs, r, te, tr, info = env.step(torch.zeros(2, 8))
    assert torch.isfinite(obs).all()
    vis = ms.make("PickCube-v1", num_envs=1, obs_mode="rgbd", robot_uids="panda_wristcam", world_factory=EmuBackendWorld)
    o, _ = vis.reset(seed=0)
    assert set(o["sensor_data"]) == {"base_camera", "hand_camera"} and o["sensor_data"]["hand_camera"]["rgb"].shape == (1, 128, 128, 3)
    with pytest.raises(NotImplementedError):
        ms.make("PickCube-v1", num_envs=1, robot_uids="xarm6_robotiq", world_factory=EmuBackendWorld)


def test_enhanced_determinism_draws_robot_noise_per_sub_scene():
    """table/scene_builder.py:85-97: with enhanced determinism every sub-scene takes its rest-pose noise from its own generator, so a
    sub-scene's reset does not depend on which other sub-scenes are reset with it."""
    env = ms.make("PickCube-v1", num_envs=3, obs_mode="state", world_factory=EmuBackendWorld, enhanced_determinism=True)
    env.reset(seed=[5, 6, 7])
    q_all = env.agent.robot.get_qpos().clone()
    env.reset(seed=[5, 6, 7])
    assert torch.allclose(env.agent.robot.get_qpos(), q_all)
    solo = ms.make("PickCube-v1", num_envs=1, obs_mode="state", world_factory=EmuBackendWorld, enhanced_determinism=True)
    solo.reset(seed=[6])
    assert torch.allclose(solo.agent.robot.get_qpos()[0], q_all[1], atol=1e-6)


def test_joint_velocity_control_modes():
    """panda.py:143-170 / pd_joint_vel.py / pd_joint_pos_vel.py: `pd_joint_vel` drives without stiffness towards a velocity target
    (action in [-1, 1] rad/s per arm joint); `pd_joint_delta_pos_vel` / `pd_joint_pos_vel` take [position part | velocity part]."""
    env = ms.make("PickCube-v1", num_envs=2, obs_mode="state", control_mode="pd_joint_vel", world_factory=EmuBackendWorld)
    env.reset(seed=0)
    assert env.action_dim == 8
    q0 = env.agent.robot.get_qpos().clone()
    a = torch.zeros(2, 8)
    a[:, 0] = 0.5                                           # +0.5 rad/s on the first joint
    a[1, 0] = -1.5                                          # clipped to -1 rad/s
    a[:, 7] = 1.0                                           # gripper stays open (position-controlled)
    for _ in range(20):                                     # 1 s
        obs, *_ = env.step(a)
    q1, qd1 = env.agent.robot.get_qpos(), env.agent.robot.get_qvel()
    assert abs(float(q1[0, 0] - q0[0, 0]) - 0.5) < 0.05 and abs(float(q1[1, 0] - q0[1, 0]) + 1.0) < 0.1
    assert abs(float(qd1[0, 0]) - 0.5) < 0.02 and abs(float(qd1[1, 0]) + 1.0) < 0.05
    # the other arm joints are velocity-held at zero: without stiffness they may creep a little under load, but do not run away
    assert (q1[:, 2:7] - q0[:, 2:7]).abs().max() < 0.05
    env.reset(options=dict(env_idx=torch.tensor([0])))
    assert torch.equal(env.scene.world.target_qvel[0, :7], torch.zeros(7)) and float(env.scene.world.target_qvel[1, 0]) == -1.0

    dpv = ms.make("PickCube-v1", num_envs=1, obs_mode="state", control_mode="pd_joint_delta_pos_vel", world_factory=EmuBackendWorld)
    dpv.reset(seed=0)
    assert dpv.action_dim == 7 + 7 + 1
    q0 = dpv.agent.robot.get_qpos().clone()
    a = torch.zeros(1, 15)
    a[0, 0], a[0, 7 + 0] = 1.0, 0.3                          # +0.1 rad target step, 0.3 rad/s feed-forward on joint 0
    dpv.step(a)
    assert float(dpv.scene.world.target_qpos[0, 0]) == pytest.approx(float(q0[0, 0]) + 0.1, abs=1e-6)
    assert float(dpv.scene.world.target_qvel[0, 0]) == pytest.approx(0.3, abs=1e-6)
    pv = ms.make("PickCube-v1", num_envs=1, obs_mode="state", control_mode="pd_joint_pos_vel", world_factory=EmuBackendWorld)
    pv.reset(seed=0)
    lo, hi = pv.single_action_space_low, pv.single_action_space_high
    assert pv.action_dim == 15 and np.allclose(lo[7:14], -1) and np.allclose(hi[7:14], 1) and lo[0] == pytest.approx(-2.8973, abs=1e-4)










