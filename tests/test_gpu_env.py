"""-m gpu: the public env API on the CUDA backend (C-ABI) -- shapes/devices (reference tests/test_gpu_envs.py:44-123),
partial reset isolation (:245-270), and the fused 5-substep control step against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_env_step_matches_oracle_and_lives_on_gpu():
    import maniskill_b200 as ms
    from oracle.oracle import OracleWorld
    n = 32
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state")
    obs, _ = env.reset(seed=11)
    assert obs.is_cuda and obs.shape == (n, 42)
    w, cm = env.scene.world, env.cm
    o = OracleWorld(cm, "f32")
    o.set_joint("qpos", w.qpos.double().cpu().numpy())
    o.set_joint("target_qpos", w.target_qpos.double().cpu().numpy())
    o.set_bodies(w.body_view().double().cpu().numpy()[:, cm.scalars["n_link"]:])
    g = torch.Generator(device=obs.device).manual_seed(0)
    for _ in range(10):
        obs, rew, term, trunc, info = env.step(2 * torch.rand((n, 8), device=obs.device, generator=g) - 1)
        o.set_joint("target_qpos", w.target_qpos.double().cpu().numpy())
        o.step(5)
    for t in (obs, rew, term, trunc, info["is_grasped"]):
        assert t.is_cuda
    ref = o.rigid_body_data()
    got = w.body_view().double().cpu().numpy()
    assert np.abs(got[..., :3] - ref[..., :3]).max() < 1e-4
    assert np.abs(w.qpos.double().cpu().numpy() - o.get_joint("qpos")).max() < 1e-4
    # obs is assembled from the same buffers
    assert torch.allclose(obs[:, :9], w.qpos[:, :9])
    env.close()


def test_partial_reset_gpu():
    import maniskill_b200 as ms
    env = ms.make("PickCube-v1", num_envs=16, obs_mode="state")
    env.reset(seed=0)
    for _ in range(3):
        obs, *_ = env.step(2 * torch.rand(16, 8, device=env.device) - 1)
    before = obs.clone()
    idx = torch.tensor([1, 5, 9], device=env.device)
    obs2, _ = env.reset(options=dict(env_idx=idx))
    keep = torch.ones(16, dtype=torch.bool, device=env.device)
    keep[idx] = False
    assert torch.allclose(obs2[keep], before[keep], atol=1e-4)
    assert not torch.allclose(obs2[idx], before[idx], atol=1e-4)
    env.close()


def test_random_rollout_stays_finite_and_bounded():
    import maniskill_b200 as ms
    env = ms.make("PickCube-v1", num_envs=256, obs_mode="state")
    venv = ms.ManiSkillVectorEnv(env)
    venv.reset(seed=5)
    g = torch.Generator(device=env.device).manual_seed(1)
    for _ in range(120):
        obs, rew, term, trunc, info = venv.step(2 * torch.rand((256, 8), device=env.device, generator=g) - 1)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert obs[:, :7].abs().max() < 4.0 and obs[:, 9:18].abs().max() < 50.0
    env.close()


def test_fused_step_matches_python_path():
    """The three-launch fused control step (b2s_pick_task_step) against the torch-op path that mirrors the reference."""
    import maniskill_b200 as ms
    n = 64
    e1 = ms.make("PickCube-v1", num_envs=n, obs_mode="state", fused=True)
    e2 = ms.make("PickCube-v1", num_envs=n, obs_mode="state", fused=False)
    assert e1._fused is not None and e2._fused is None
    o1, _ = e1.reset(seed=21)
    o2, _ = e2.reset(seed=21)
    assert torch.allclose(o1, o2, atol=1e-6)
    g = torch.Generator(device=e1.device).manual_seed(3)
    # drive the gripper closed around the cube in a few envs so that is_grasped is exercised
    for i in range(25):
        a = 2 * torch.rand((n, 8), device=e1.device, generator=g) - 1
        a[:, 7] = -1.0 if i > 5 else 1.0
        o1, r1, t1, tr1, i1 = e1.step(a)
        o2, r2, t2, tr2, i2 = e2.step(a)
        assert torch.allclose(o1, o2, atol=2e-5), (i, (o1 - o2).abs().max())
        assert torch.allclose(r1, r2, atol=2e-5)
        assert torch.equal(t1, t2)
        for k in ("success", "is_obj_placed", "is_robot_static", "is_grasped"):
            assert torch.equal(i1[k], i2[k]), k
        assert torch.equal(i1["elapsed_steps"], i2["elapsed_steps"])
    e1.close()
    e2.close()


def test_fused_step_serves_visual_obs_modes(monkeypatch):
    """The fused control step also runs under the visual observation modes (default; B2S_FUSED_VISUAL=0 turns it off): same
    observation dict (state, sensor parameters, sensor data) as the torch path."""
    import maniskill_b200 as ms
    n = 16
    monkeypatch.setenv("B2S_FUSED_VISUAL", "1")
    e1 = ms.make("PickCube-v1", num_envs=n, obs_mode="state+rgb+depth", fused=True)
    e2 = ms.make("PickCube-v1", num_envs=n, obs_mode="state+rgb+depth", fused=False)
    assert e1._fused is not None and e2._fused is None
    e1.reset(seed=5)
    e2.reset(seed=5)
    g = torch.Generator(device=e1.device).manual_seed(1)
    for i in range(12):
        a = 2 * torch.rand((n, 8), device=e1.device, generator=g) - 1
        o1, r1, t1, _, i1 = e1.step(a)
        o2, r2, t2, _, i2 = e2.step(a)
        # a visual mode with the `state` flag carries one flat state vector in place of agent / extra (sapien_env.py:540-544)
        assert set(o1.keys()) == set(o2.keys()) == {"state", "sensor_param", "sensor_data"}
        assert o1["state"].shape == (n, 42) and torch.allclose(o1["state"], o2["state"], atol=2e-5)
        for k in ("rgb", "depth"):
            d = (o1["sensor_data"]["base_camera"][k].int() - o2["sensor_data"]["base_camera"][k].int()).abs()
            assert (d > 1).float().mean() < 1e-3, k  # the two envs' physics agree to ~1e-5: a handful of edge pixels may flip
        for k in o1["sensor_param"]["base_camera"]:
            assert torch.allclose(o1["sensor_param"]["base_camera"][k], o2["sensor_param"]["base_camera"][k], atol=1e-5)
        assert torch.allclose(r1, r2, atol=2e-5) and torch.equal(t1, t2)
    e1.close()
    e2.close()


@pytest.mark.parametrize("task,steps", [("PegInsertionSide-v1", 8), ("OpenCabinetDrawer-v1", 6)])
def test_other_tasks_match_oracle(task, steps):
    """Heterogeneous per-env geometry (peg) and the two-articulation / 17-dof configuration (Fetch + cabinet, the large
    kernel instantiation) against the CPU oracle started from the same state."""
    from scenarios import run_task_vs_oracle
    err_q, err_p, overflow = run_task_vs_oracle(task, steps, 16)
    assert err_q < 1e-4 and err_p < 1e-4, (err_q, err_p)
    assert overflow == 0


def test_ee_target_controller_on_gpu():
    """pd_ee_target_delta_pose through the torch path on the device (batched FK / Jacobian / damped least squares on CUDA): the tcp
    reaches the accumulated target (CPU twin: tests/test_ee_controller.py)."""
    import maniskill_b200 as ms
    n = 32
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode="pd_ee_target_delta_pose")
    assert env._fused is None and env.action_dim == 7
    env.reset(seed=11)
    p0 = env.agent.tcp.pose.p.clone()
    a = torch.zeros((n, 7), device=env.device)
    a[:, 0], a[:, 2], a[:, -1] = 0.5, -0.4, 1.0
    for _ in range(3):
        env.step(a)
    hold = torch.zeros((n, 7), device=env.device)
    hold[:, -1] = 1.0
    for _ in range(12):
        obs, rew, term, trunc, info = env.step(hold)
    moved = env.agent.tcp.pose.p - p0
    want = torch.tensor([0.15, 0.0, -0.12], device=env.device)
    assert (moved - want).abs().max() < 0.01, moved
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    env.close()

