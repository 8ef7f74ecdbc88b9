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
        ms.make("PickCube-v1", num_envs=1, obs_mode="pointcloud", world_factory=EmuBackendWorld)
    with pytest.raises(NotImplementedError):
        ms.make("PickCube-v1", num_envs=1, control_mode="pd_ee_delta_pose", world_factory=EmuBackendWorld)
