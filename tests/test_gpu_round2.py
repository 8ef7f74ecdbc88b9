"""-m gpu: round-2 parity and behaviour checks through the C-ABI.

* PickCube-v1 at the BENCHMARKED size (4096 sub-scenes, many blocks): 64 sampled sub-scenes against the CPU oracle after 100
  substeps, no capacity overflow (parity ran at 64 envs only)
* every registered task x 1 control step against the oracle, overflow == 0
* the device-side auto-reset (b2s_pick_task_autoreset) against the python flow of `ManiSkillVectorEnv` (the mirror of
  mani_skill/vector/wrappers/gymnasium.py:127-184)
* masked re-rendering, `final_observation` pictures (they used to alias the re-rendered targets)
"""
import numpy as np
import pytest
import torch

from scenarios import pick_cube_random_actions, rel_err_vec, run_task_vs_oracle

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def test_pick_cube_4096_envs_sampled_vs_oracle_no_overflow():
    """The benchmark's size: 4096 sub-scenes through CUDA, the same scenario for 64 sampled sub-scenes through the oracle."""
    from maniskill_b200.backend import BUF_ALL, World
    from maniskill_b200.scenes import pick_cube_scene
    from oracle.oracle import OracleWorld
    N, n_sub = 4096, 100
    cm = pick_cube_scene(N).compile()
    q0, cube, deltas, grip = pick_cube_random_actions(N, n_sub, seed=1)
    sample = np.sort(np.random.RandomState(0).choice(N, 64, replace=False))
    sample[:4] = [0, 31, 32, N - 1]  # block / warp boundaries
    sample = np.unique(sample)
    cube_fb, n_link = cm.actor_fb["cube"], cm.scalars["n_link"]
    # CUDA, all sub-scenes
    w = World(cm)
    dev = w.device
    w.qpos[:] = torch.tensor(q0, dtype=torch.float32, device=dev)
    w.target_qpos[:] = w.qpos
    w.body_view()[:, n_link + cube_fb] = torch.tensor(cube, dtype=torch.float32, device=dev)
    w.apply()
    for s in range(n_sub):
        if s % 5 == 0:
            w.fetch(4)
            tq = w.qpos.double() + torch.tensor(deltas[s // 5], device=dev)
            tq[:, 7:] = torch.tensor(grip[s // 5], device=dev)
            w.target_qpos[:] = tq.float()
            w.apply(32)
        w.step(1, 0)
    w.fetch(BUF_ALL)
    torch.cuda.synchronize()
    got_q, got_qd = w.qpos.double().cpu().numpy()[sample], w.qvel.double().cpu().numpy()[sample]
    got_body = w.body_view().double().cpu().numpy()[sample]
    assert int(w.overflow_flag.item()) == 0, w.overflow_reasons()
    assert np.isfinite(w.body_view().cpu().numpy()).all()
    # oracle, the sampled sub-scenes with the very same inputs
    cm_s = pick_cube_scene(len(sample)).compile()
    o = OracleWorld(cm_s, "f32")
    o.set_joint("qpos", q0[sample])
    o.set_joint("target_qpos", q0[sample])
    b = o.get_bodies()
    b[:, cube_fb] = cube[sample]
    o.set_bodies(b)
    for s in range(n_sub):
        if s % 5 == 0:
            tq = o.get_joint("qpos") + deltas[s // 5][sample]
            tq[:, 7:] = grip[s // 5][sample]
            o.set_joint("target_qpos", tq)
        o.step(1)
    ref_body = o.rigid_body_data()
    e_q = rel_err_vec(got_q, o.get_joint("qpos"), 0.1)
    e_qd = rel_err_vec(got_qd, o.get_joint("qvel"), 0.1)
    e_pos = rel_err_vec(got_body[..., :3], ref_body[..., :3], 0.1)
    e_quat = rel_err_vec(got_body[..., 3:7], ref_body[..., 3:7], 1.0)
    print("4096 envs, 64 sampled: rel err q %.3g qd %.3g pos %.3g quat %.3g" % (e_q, e_qd, e_pos, e_quat))
    assert e_q < REL_TOL and e_qd < REL_TOL and e_pos < REL_TOL and e_quat < REL_TOL
    w.close()


ALL_TASKS = ["PickCube-v1", "PegInsertionSide-v1", "OpenCabinetDrawer-v1"]


def _registered_tasks():
    import maniskill_b200 as ms
    reg = getattr(ms, "REGISTERED_ENVS", None) or getattr(ms.envs, "REGISTERED_ENVS", None)
    return sorted(reg.keys()) if reg else ALL_TASKS


@pytest.mark.parametrize("task", _registered_tasks())
def test_every_task_one_control_step_vs_oracle(task):
    err_q, err_p, overflow = run_task_vs_oracle(task, steps=1, n=16)
    print(task, "|dq| %.3g |dp| %.3g overflow %d" % (err_q, err_p, overflow))
    assert overflow == 0
    assert err_q < 1e-4 and err_p < 1e-4


def _pick_envs(n, obs_mode="state", **kw):
    import maniskill_b200 as ms
    a = ms.ManiSkillVectorEnv(ms.make("PickCube-v1", num_envs=n, obs_mode=obs_mode), **kw)
    b = ms.ManiSkillVectorEnv(ms.make("PickCube-v1", num_envs=n, obs_mode=obs_mode), device_autoreset=False, **kw)
    assert a._device_autoreset and not b._device_autoreset
    return a, b


def test_device_autoreset_matches_python_flow_until_the_reset_and_resets_correctly():
    n = 96
    dev_env, py_env = _pick_envs(n)
    o1, _ = dev_env.reset(seed=11)
    o2, _ = py_env.reset(seed=11)
    assert torch.equal(o1, o2)
    g = torch.Generator(device=o1.device).manual_seed(5)
    for t in range(50):
        a = 2 * torch.rand((n, 8), device=o1.device, generator=g) - 1
        o1, r1, te1, tr1, i1 = dev_env.step(a)
        o2, r2, te2, tr2, i2 = py_env.step(a)
        same = ~(te2 | tr2) if t < 49 else None
        assert torch.equal(r1, r2) and torch.equal(te1, te2) and torch.equal(tr1, tr2), t
        if t < 49:
            rows = torch.nonzero(~(te2 | tr2))[:, 0]
            assert torch.equal(o1[rows], o2[rows]), t
            for k in ("success", "is_grasped", "is_obj_placed", "is_robot_static", "elapsed_steps"):
                assert torch.equal(i1[k][rows], i2[k][rows]), (t, k)
        assert not i1["_final_info"][~(te1 | tr1)].any()
    # step 50: every sub-scene hits the time limit
    assert bool(tr1.all()) and bool(i1["_final_info"].all())
    assert torch.equal(i1["final_observation"], i2["final_observation"])
    assert torch.equal(i1["final_info"]["success"], i2["final_info"]["success"])
    assert torch.equal(i1["final_info"]["elapsed_steps"], i2["final_info"]["elapsed_steps"])
    env = dev_env.base_env
    assert int(env.elapsed_steps.abs().sum()) == 0
    nd = 9
    q, qd = o1[:, :nd], o1[:, nd:2 * nd]
    from maniskill_b200.envs.tabletop import REST_QPOS
    rest = torch.tensor(REST_QPOS["panda"], dtype=torch.float32, device=o1.device)
    assert float(qd.abs().max()) == 0.0
    assert float((q[:, :7] - rest[:7]).abs().max()) < 0.12 and float((q[:, :7] - rest[:7]).std()) == pytest.approx(0.02, rel=0.25)
    assert torch.equal(q[:, 7:], torch.full((n, 2), 0.04, device=o1.device))
    obj, goal = o1[:, 2 * nd + 11:2 * nd + 18], o1[:, 2 * nd + 8:2 * nd + 11]
    assert float(obj[:, :2].abs().max()) <= 0.1 + 1e-6 and torch.allclose(obj[:, 2], torch.full((n,), 0.02, device=o1.device))
    assert float(obj[:, :2].std()) > 0.03                      # spread over the spawn square, not a constant
    assert torch.allclose(obj[:, 3:].norm(dim=1), torch.ones(n, device=o1.device), atol=1e-5) and float(obj[:, 4:6].abs().max()) == 0.0
    assert float(goal[:, :2].abs().max()) <= 0.1 + 1e-6 and float(goal[:, 2].min()) >= 0.02 - 1e-6 and float(goal[:, 2].max()) <= 0.32 + 1e-6
    # the world really is in that state: a python-side fetch agrees with the observation row
    w = env.scene.world
    w.fetch()
    assert torch.allclose(w.qpos[:, :nd], q) and torch.equal(w.target_qpos[:, :nd], w.qpos[:, :nd])
    # and the episode goes on from there
    for t in range(3):
        o1, r1, te1, tr1, i1 = dev_env.step(2 * torch.rand((n, 8), device=o1.device, generator=g) - 1)
    assert torch.isfinite(o1).all() and int(env.elapsed_steps.min()) == 3
    assert int(w.overflow_flag.item()) == 0
    dev_env.close(); py_env.close()


def test_device_autoreset_partial_and_ignore_terminations():
    """Only finished sub-scenes are touched: a sub-scene forced over the time limit resets, its neighbours continue untouched."""
    import maniskill_b200 as ms
    n = 64
    venv = ms.ManiSkillVectorEnv(ms.make("PickCube-v1", num_envs=n, obs_mode="state"), ignore_terminations=True)
    assert venv._device_autoreset
    obs, _ = venv.reset(seed=3)
    env = venv.base_env
    env._elapsed_steps[5] = 49
    env._elapsed_steps[40] = 49
    a = torch.zeros((n, 8), device=obs.device)
    o, r, te, tr, info = venv.step(a)
    done = info["_final_info"]
    assert done.nonzero()[:, 0].tolist() == [5, 40] and tr.nonzero()[:, 0].tolist() == [5, 40] and not te.any()
    assert env.elapsed_steps[5].item() == 0 and env.elapsed_steps[6].item() == 1
    assert info["final_info"]["elapsed_steps"][5].item() == 50
    venv.close()


def test_masked_render_and_final_observation_pictures():
    import maniskill_b200 as ms
    from oracle import raster
    n = 8
    venv = ms.ManiSkillVectorEnv(ms.make("PickCube-v1", num_envs=n, obs_mode="state+rgb+depth+segmentation"))
    assert venv._device_autoreset
    obs, _ = venv.reset(seed=2)
    env = venv.base_env
    g = torch.Generator(device=env.device).manual_seed(1)
    for _ in range(3):
        obs, *_ = venv.step(2 * torch.rand((n, 8), device=env.device, generator=g) - 1)
    env._elapsed_steps[2] = 49
    before = {k: v.clone() for k, v in obs["sensor_data"]["base_camera"].items()}
    obs, r, te, tr, info = venv.step(torch.zeros((n, 8), device=env.device))
    done = info["_final_observation"]
    assert done.nonzero()[:, 0].tolist() == [2]
    torch.cuda.synchronize()
    # pictures of every sub-scene equal the oracle's rendering of the CURRENT state (sub-scene 2: the state after its reset)
    body = env.scene.world.body_view().cpu().numpy()
    (color, posseg), = raster.render(env._sensors.visuals, env._sensors.cams, body)
    sd = obs["sensor_data"]["base_camera"]
    assert np.array_equal(sd["segmentation"][..., 0].cpu().numpy(), posseg[..., 3])
    assert np.array_equal(sd["depth"][..., 0].cpu().numpy(), -posseg[..., 2])
    assert np.abs(sd["rgb"].cpu().numpy().astype(int) - color[..., :3].astype(int)).max() <= 1
    # final_observation of the finished sub-scene shows the finished state, not the re-rendered one
    fsd = info["final_observation"]["sensor_data"]["base_camera"]
    assert not torch.equal(fsd["rgb"][2], sd["rgb"][2])
    assert fsd["rgb"].data_ptr() != sd["rgb"].data_ptr()
    assert info["final_observation"]["state"].shape == obs["state"].shape
    venv.close()


def test_python_flow_final_observation_is_not_aliased():
    """with dict observations `final_observation` must be a deep copy (gymnasium.py:165 torch_clone_dict)."""
    import maniskill_b200 as ms
    n = 4
    venv = ms.ManiSkillVectorEnv(ms.make("PickCube-v1", num_envs=n, obs_mode="rgb"), device_autoreset=False)
    obs, _ = venv.reset(seed=2)
    venv.base_env._elapsed_steps[:] = 49
    obs, r, te, tr, info = venv.step(torch.zeros((n, 8), device=venv.device))
    assert "final_observation" in info
    a, b = info["final_observation"]["sensor_data"]["base_camera"]["rgb"], obs["sensor_data"]["base_camera"]["rgb"]
    assert a.data_ptr() != b.data_ptr() and not torch.equal(a, b)
    venv.close()


def test_fused_step_rejects_wrong_action_shape_and_applies_pending_setters():
    import maniskill_b200 as ms
    n = 8
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state")
    env.reset(seed=0)
    with pytest.raises(ValueError):
        env.step(torch.zeros((n, 5), device=env.device))
    with pytest.raises(ValueError):
        env.step(torch.zeros((3, 8), device=env.device))
    # a single action is broadcast (the torch path does the same)
    o1, *_ = env.step(torch.zeros(8, device=env.device))
    # set_qpos before a fused step is not discarded
    q = env.agent.robot.get_qpos().clone()
    q[:, 0] += 0.3
    env.agent.robot.set_qpos(q)
    o2, *_ = env.step(torch.zeros((n, 8), device=env.device))
    assert float((o2[:, 0] - q[:, 0]).abs().max()) < 0.05
    # fresh tensors every step
    o3, r3, *_ = env.step(torch.zeros((n, 8), device=env.device))
    assert o3.data_ptr() != o2.data_ptr()
    env.close()


def test_world_on_a_device_that_is_not_current():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import maniskill_b200 as ms
    torch.cuda.set_device(0)
    env = ms.make("PickCube-v1", num_envs=16, obs_mode="state", device="cuda:1")
    obs, _ = env.reset(seed=0)
    for _ in range(3):
        obs, *_ = env.step(torch.zeros((16, 8), device="cuda:1"))
    assert obs.device.index == 1 and torch.isfinite(obs).all() and torch.cuda.current_device() == 0
    env.close()


@pytest.mark.parametrize("task,obs_mode", [("PegInsertionSide-v1", "state"), ("PegInsertionSide-v1", "rgbd"), ("OpenCabinetDrawer-v1", "state")])
def test_graphed_epilogue_equals_the_eager_one(task, obs_mode, monkeypatch):
    """Tasks without a hand-written fused epilogue replay evaluate / observation / reward as one captured CUDA graph
    (base_env._GraphedEpilogue).  Same kernels in the same order as the eager code: results must be identical, across a partial reset and
    through the vector wrapper's auto-reset."""
    import torch
    import maniskill_b200 as ms

    def rollout(graphed):
        monkeypatch.setenv("B2S_GRAPH_EPILOGUE", "1" if graphed else "0")
        torch.manual_seed(0)   # the episode initialisers draw from torch's global generator
        env = ms.make(task, num_envs=32, obs_mode=obs_mode, device="cuda:0")
        assert (env._epilogue_runner is not None) == graphed
        venv = ms.ManiSkillVectorEnv(env, max_episode_steps=6)
        obs, _ = venv.reset(seed=5)
        gen = torch.Generator(device="cuda:0")
        gen.manual_seed(11)
        out = []
        for i in range(9):
            a = 2 * torch.rand((32, env.action_dim), device="cuda:0", generator=gen) - 1
            obs, rew, term, trunc, info = venv.step(a)
            if i == 3:
                obs, _ = venv.reset(options=dict(env_idx=torch.tensor([1, 7, 30], device="cuda:0")))
            state = obs["state"] if isinstance(obs, dict) and "state" in obs else obs
            rec = dict(rew=rew.clone(), term=term.clone(), trunc=trunc.clone(), success=info["success"].clone())
            if isinstance(state, torch.Tensor):
                rec["state"] = state.clone()
            if isinstance(obs, dict) and "sensor_data" in obs:
                rec["rgb"] = obs["sensor_data"]["base_camera"]["rgb"].clone()
                rec["tcp"] = obs["extra"]["tcp_pose"].clone() if "extra" in obs else None
            out.append(rec)
        if graphed:
            assert env._epilogue_runner.graph is not None, "the epilogue was never captured"
        torch.cuda.synchronize()
        assert int(env.scene.world.overflow_flag.item()) == 0
        venv.close()
        return out

    a, b = rollout(False), rollout(True)
    for i, (x, y) in enumerate(zip(a, b)):
        for k in x:
            if x[k] is None:
                continue
            assert torch.equal(x[k], y[k]), (task, obs_mode, i, k)


def test_record_and_replay_on_the_device(tmp_path):
    """RecordEpisode / replay_trajectory (mani_skill/utils/wrappers/record.py, mani_skill/trajectory/replay_trajectory.py) over the CUDA
    world: 8 sub-scenes recorded through the fused control step with device tensors, partial flush after a partial reset, then replayed
    from seed + actions and from the stored environment states on a fresh world."""
    import torch
    import maniskill_b200 as ms
    from maniskill_b200.trajectory import RecordEpisode, load_trajectories, replay_trajectory
    env = ms.make("PickCube-v1", num_envs=1, obs_mode="state", device="cuda:0")
    rec = RecordEpisode(env, str(tmp_path))
    g = torch.Generator(device="cuda:0").manual_seed(5)
    for seed in (11, 12):
        rec.reset(seed=seed)
        for _ in range(6):
            rec.step(2 * torch.rand(1, 8, generator=g, device="cuda:0") - 1)
    rec.close()
    meta, trajs = load_trajectories(str(tmp_path / "trajectory"))
    assert [e["episode_seed"] for e in meta["episodes"]] == [11, 12] and trajs["traj_0"]["actions"].shape == (6, 8)
    for mode in ("seed_and_actions", "env_states"):
        env2 = ms.make("PickCube-v1", num_envs=1, obs_mode="state", device="cuda:0")
        if mode != "seed_and_actions":
            env2.reset(seed=999)
        res = replay_trajectory(env2, str(tmp_path / "trajectory"), use_env_states=mode == "env_states")
        assert len(res) == 2
        for r in res:
            assert r["final_state_error"] < 1e-5 and r["success"] == r["recorded_success"], (mode, r)
        env2.close()
    # many sub-scenes, partial reset -> partial flush
    envn = ms.make("PickCube-v1", num_envs=8, obs_mode="state", device="cuda:0")
    recn = RecordEpisode(envn, str(tmp_path / "many"))
    recn.reset(seed=3)
    for _ in range(4):
        recn.step(2 * torch.rand(8, 8, generator=g, device="cuda:0") - 1)
    recn.reset(options=dict(env_idx=torch.tensor([2, 5], device="cuda:0")))
    for _ in range(2):
        recn.step(2 * torch.rand(8, 8, generator=g, device="cuda:0") - 1)
    recn.close()
    meta, trajs = load_trajectories(str(tmp_path / "many" / "trajectory"))
    lens = sorted(t["actions"].shape[0] for t in trajs.values())
    assert lens == [2, 2, 4, 4, 6, 6, 6, 6, 6, 6], lens


@pytest.mark.parametrize("task", ["PickCube-v1", "OpenCabinetDrawer-v1"])
def test_group_dynamics_kernel_equals_the_lane_kernel(task, monkeypatch):
    """The dynamics half of kin with eight lanes per sub-scene (kin_dyn_kernel: articulated inertias in shared memory, lane r owns row r of
    the 6x6 sweeps, shuffle reductions inside the group, one M~^-1 column per lane; B2S_KIN_GROUP=1) against the one-lane kernel: same
    arithmetic, different summation order in three reductions per joint."""
    import torch
    import maniskill_b200 as ms

    def rollout(group):
        monkeypatch.setenv("B2S_KIN_GROUP", "1" if group else "0")
        torch.manual_seed(0)
        env = ms.make(task, num_envs=64, obs_mode="state", device="cuda:0")
        env.reset(seed=9)
        gen = torch.Generator(device="cuda:0")
        gen.manual_seed(2)
        for _ in range(6):
            env.step(2 * torch.rand((64, env.action_dim), device="cuda:0", generator=gen) - 1)
        w = env.scene.world
        torch.cuda.synchronize()
        out = (w.qpos.clone(), w.qvel.clone(), w.body_view().clone(), int(w.overflow_flag.item()))
        env.close()
        return out

    a, b = rollout(False), rollout(True)
    assert a[3] == 0 and b[3] == 0
    assert torch.isfinite(b[2]).all()
    # 30 substeps of a contact-rich system amplify the last-bit differences of the reductions; the bar of the oracle parity tests is 1e-4
    assert (a[0] - b[0]).abs().max() < 1e-4 and (a[2][..., :7] - b[2][..., :7]).abs().max() < 1e-4
    assert (a[1] - b[1]).abs().max() < 2e-2 * max(1.0, float(a[1].abs().max()))


@pytest.mark.parametrize("task,mode", [("PickCube-v1", "pd_ee_delta_pose"), ("PickCube-v1", "pd_ee_delta_pos"), ("PegInsertionSide-v1", "pd_ee_delta_pose")])
def test_ik_kernel_matches_the_torch_step(task, mode):
    """b2s_ik_step (one damped least-squares step per sub-scene in one kernel: chain FK, Jacobian, 6 x 6 Cholesky) against the batched
    torch restatement of the reference's GPU branch (mani_skill/agents/controllers/utils/kinematics.py:243-260: (J^T J + 1e-4 I) dq = J^T d).
    fp32, damping 1e-4: tolerance 2e-4 rad on the joint targets for end-effector displacements of a few centimetres."""
    import torch
    import maniskill_b200 as ms
    n = 96
    env = ms.make(task, num_envs=n, obs_mode="state", control_mode=mode, device="cuda:0")
    env.reset(seed=4)
    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(8)
    for _ in range(3):   # move away from the rest pose
        env.step(2 * torch.rand((n, env.action_dim), device="cuda:0", generator=gen) - 1)
    ctrl = env.agent.controller.controllers["arm"]
    kin = ctrl.kinematics
    assert kin._ik_args is not None, "the CUDA world must provide the IK entry point"
    delta = (2 * torch.rand((n, 6), device="cuda:0", generator=gen) - 1) * torch.tensor([0.05, 0.05, 0.05, 0.2, 0.2, 0.2], device="cuda:0")
    q0 = env.agent.robot.get_qpos()
    cfg = dict(type="levenberg_marquardt", alpha=1.0)
    got = kin.compute_ik(delta, q0, cfg)
    args, kin._ik_args = kin._ik_args, None
    try:
        ref = kin.compute_ik(delta, q0, cfg)
    finally:
        kin._ik_args = args
    torch.cuda.synchronize()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert (got - ref).abs().max() < 2e-4, float((got - ref).abs().max())
    # the step moves the end effector by the requested translation (first order): FK of the new targets
    q1 = q0.clone()
    cols = kin.chain_dof_idx[kin.qmask]
    q1[:, cols] = got
    p0, _ = kin.fk(q0)
    p1, _ = kin.fk(q1)
    assert ((p1 - p0) - delta[:, :3]).abs().max() < 0.02
    env.close()
