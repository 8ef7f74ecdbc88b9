"""The UNMODIFIED reference package on the `sapien` shim (maniskill_b200/compat) -- SURVEY.md section 7 step 2's gate and VERDICT r1 item 4.

`/root/reference/mani_skill` is imported byte for byte; `sapien`, `gymnasium`, `dacite`, `transforms3d`, ... resolve to
maniskill_b200/compat/site (none of them is installed here).  The world the shim creates at `px.gpu_init()` is the host emulation of
the device code (tests/emu) -- this box has no GPU -- and the only thing patched in the reference is the torch device its
`parse_sim_and_render_backend` returns (cuda -> cpu), so that its tensors live where the emulated buffers do.

Checked: `gym.make("PickCube-v1", num_envs=16)` builds through the reference's own builders / URDF loader / agent / controllers; the
reference's OWN tests -- tests/test_gpu_envs.py::test_partial_resets, tests/test_sim_state.py::test_raw_sim_states (state width 70) --
pass unmodified; a rollout agrees with this repo's mirror of the task to 1e-5; visual observation modes render.
Skipped where /root/reference does not exist (the GPU box).
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mani_skill")), reason="needs the reference checkout at /root/reference")


@pytest.fixture(scope="module")
def reference():
    """Installs the shim + the emulated world, imports the reference, patches its device choice.  Yields the gymnasium module."""
    import maniskill_b200.compat as compat
    from emu_world import EmuBackendWorld
    compat.install()
    compat.WORLD_FACTORY = lambda cm, dev: EmuBackendWorld(cm)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "MS_ASSET_DIR" not in os.environ:
        # OpenCabinetDrawer-v1 reads PartNet-Mobility cabinets from $MS_ASSET_DIR/data (a download that is not available): stand-in URDFs of the same kind
        # (tools/make_standin_partnet.py); the variable is read when mani_skill is imported
        import tempfile
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
        import make_standin_partnet
        assets = tempfile.mkdtemp(prefix="b200sim_ms_assets_")
        make_standin_partnet.main(os.path.join(REF, "mani_skill", "assets", "partnet_mobility", "meta"), assets)
        os.environ["MS_ASSET_DIR"] = assets
    import gymnasium as gym
    import mani_skill.envs  # noqa: F401  (registers every task: all task modules, scene builders and robots import)
    import mani_skill.envs.sapien_env as SE
    import mani_skill.envs.utils.system.backend as B
    orig = B.parse_sim_and_render_backend

    def parse(sim_backend, render_backend):
        info = orig(sim_backend, render_backend)
        info.device = torch.device("cpu")
        return info

    SE.parse_sim_and_render_backend = parse
    # the emulated world IS the GPU simulation: the reference's controllers must take their GPU kinematics path (batched torch FK / Jacobian,
    # mani_skill/agents/controllers/utils/kinematics.py:88-93 picks it from device.type == "cuda"), not pinocchio
    import mani_skill.agents.controllers.utils.kinematics as K
    setup_cpu = K.Kinematics._setup_cpu
    K.Kinematics._setup_cpu = K.Kinematics._setup_gpu
    sync = torch.cuda.synchronize
    if not torch.cuda.is_available():   # sapien_env.py:624 synchronises the device after rendering; there is none on this box
        torch.cuda.synchronize = lambda *a, **k: None
    yield gym
    torch.cuda.synchronize = sync
    K.Kinematics._setup_cpu = setup_cpu
    SE.parse_sim_and_render_backend = orig
    compat.WORLD_FACTORY = None


def test_the_shim_is_what_got_imported(reference):
    import dacite
    import gymnasium
    import sapien
    import transforms3d
    site = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maniskill_b200", "compat", "site")
    for m in (sapien, gymnasium, dacite, transforms3d):
        assert m.__file__.startswith(site), m.__file__
    import mani_skill
    assert mani_skill.__file__.startswith(REF)


def _run_reference_test(module_file, fn_name, *args):
    """Executes one test function of /root/reference/tests/<module_file> as it is (the module does `from tests.utils import ...`)."""
    spec = importlib.util.spec_from_file_location("ref_" + module_file[:-3], os.path.join(REF, "tests", module_file))
    ref_tests = importlib.util.spec_from_file_location("tests", os.path.join(REF, "tests", "__init__.py"), submodule_search_locations=[os.path.join(REF, "tests")])
    saved = {k: v for k, v in sys.modules.items() if k == "tests" or k.startswith("tests.")}
    try:
        pkg = importlib.util.module_from_spec(ref_tests)
        sys.modules["tests"] = pkg
        ref_tests.loader.exec_module(pkg)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        torch.manual_seed(0)   # the tests reset without a seed: episode layouts then come from torch's global generator (test_timelimits needs
        np.random.seed(0)      # "no sub-scene succeeds by chance within 50 idle steps", which one layout in a dozen violates)
        getattr(mod, fn_name)(*args)
    finally:
        for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("fn,args", [
    ("test_env_control_modes", ("PickCube-v1", "pd_joint_delta_pos")), ("test_env_control_modes", ("PickCube-v1", "pd_joint_pos")),
    ("test_env_control_modes", ("PickCube-v1", "pd_ee_delta_pose")), ("test_env_control_modes", ("PickCube-v1", "pd_ee_delta_pos")),
    ("test_env_control_modes", ("StackCube-v1", "pd_ee_delta_pose")),
    ("test_robots", ("PickCube-v1", "panda")), ("test_multi_agent", ("TwoRobotPickCube-v1",)), ("test_timelimits", ()), ("test_hidden_objs", ("PickCube-v1",)),
    ("test_wrappers.py:test_multi_agent_flatten_action_space_gpu", ("TwoRobotStackCube-v1",)),
], ids=lambda v: "-".join(v) if isinstance(v, tuple) else str(v))
def test_reference_own_gpu_env_tests(reference, fn, args):
    """/root/reference/tests/test_gpu_envs.py: control modes (the end-effector ones through the reference's GPU kinematics over the
    pytorch_kinematics stand-in), robots, multi-agent, time limits, hidden objects; tests/test_wrappers.py: the flattened multi-agent action
    space -- 16 sub-scenes each, executed as they are.
    (`test_envs_obs_modes` asserts `device == cuda:0` and can only run where a GPU and the reference are on the same machine.)"""
    module_file, _, fn = fn.rpartition(":")
    _run_reference_test(module_file or "test_gpu_envs.py", fn, *args)


@pytest.mark.parametrize("fn,args", [
    ("test_sim_state.py:test_raw_heterogeneous_actor_sim_states", ()), ("test_sim_state.py:test_raw_heterogeneous_articulations_sim_states", ()),
    ("test_gpu_envs.py:test_env_control_modes", ("PegInsertionSide-v1", "pd_joint_pos")), ("test_gpu_envs.py:test_env_control_modes", ("PegInsertionSide-v1", "pd_ee_delta_pose")),
    ("test_gpu_envs.py:test_env_control_modes", ("StackCube-v1", "pd_joint_delta_pos")), ("test_gpu_envs.py:test_env_control_modes", ("StackCube-v1", "pd_ee_delta_pos")),
    ("test_gpu_envs.py:test_robots", ("StackCube-v1", "panda")), ("test_gpu_envs.py:test_multi_agent", ("TwoRobotStackCube-v1",)),
    ("structs/test_actor.py:test_actor_pose_gpu", ()), ("structs/test_link.py:test_link_pose_gpu", ()),
    ("structs/test_pose.py:test_pose_creation", ()), ("structs/test_pose.py:test_pose_create_with_p", ()), ("structs/test_pose.py:test_pose_create_with_q", ()),
    ("structs/test_pose.py:test_pose_to_sapien_pose", ()), ("structs/test_pose.py:test_pose_mult", ()), ("structs/test_pose.py:test_pose_inv", ()),
    ("structs/test_pose.py:test_pose_transformation_matrix", ()),
], ids=lambda v: "-".join(v) if isinstance(v, tuple) else str(v))
def test_reference_own_tests_second_batch(reference, fn, args):
    """More of /root/reference/tests executed as they are: the state get / set round trip of PegInsertionSide-v1 (sub-scenes with different peg and
    hole geometry, state width 13 * 3 + 13 + 9 * 2; tests/test_sim_state.py:40-71) and of OpenCabinetDrawer-v1 (Fetch + one cabinet per sub-scene merged into one
    view, state width 13 + 13 + 2 * max_dof + 13 + 15 * 2; :73-103; stand-in cabinet assets, see the fixture), further task x control-mode / robot / multi-agent cases of
    tests/test_gpu_envs.py, and tests/structs/ (Actor / Link pose setters on the GPU buffers, the Pose struct over sapien.Pose)."""
    module_file, _, fn = fn.rpartition(":")
    _run_reference_test(module_file, fn, *args)


@pytest.mark.parametrize("env_id,obs_mode", [("PickCube-v1", m) for m in ("state_dict", "state", "rgb", "rgb+depth+segmentation", "pointcloud", "depth+state", "state+rgb+segmentation")]
                         + [("StackCube-v1", m) for m in ("state_dict", "rgb+depth+segmentation", "pointcloud")]
                         + [("PegInsertionSide-v1", m) for m in ("state", "rgb", "rgb+depth+segmentation", "pointcloud", "depth+state", "state+rgb+segmentation")])
def test_reference_own_test_envs_obs_modes(reference, env_id, obs_mode):
    """/root/reference/tests/test_gpu_envs.py:44-121 (`test_envs_obs_modes`: tensor types, camera texture shapes / dtypes, sensor parameters and point clouds
    of every observation mode).  Its helper asserts `x.device == torch.device("cuda:0")`; this box has no GPU, so the ONE literal "cuda:0" of the function's
    source is replaced by "cpu" when it is loaded -- everything else runs as written (OBS_MODES of tests/utils.py; PegInsertionSide has two cameras)."""
    import inspect
    import textwrap
    holder = {}

    def grab(mod_fn):
        holder["fn"] = mod_fn

    # load the module the same way as every other reference test, then re-compile the one function with the device literal changed
    spec = importlib.util.spec_from_file_location("ref_test_gpu_envs_obs", os.path.join(REF, "tests", "test_gpu_envs.py"))
    ref_tests = importlib.util.spec_from_file_location("tests", os.path.join(REF, "tests", "__init__.py"), submodule_search_locations=[os.path.join(REF, "tests")])
    saved = {k: v for k, v in sys.modules.items() if k == "tests" or k.startswith("tests.")}
    try:
        pkg = importlib.util.module_from_spec(ref_tests)
        sys.modules["tests"] = pkg
        ref_tests.loader.exec_module(pkg)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        src = textwrap.dedent(inspect.getsource(mod.test_envs_obs_modes))
        src = src[src.index("def test_envs_obs_modes"):]
        assert src.count('"cuda:0"') == 1
        exec(compile(src.replace('"cuda:0"', '"cpu"'), "test_gpu_envs.py::test_envs_obs_modes", "exec"), mod.__dict__)
        torch.manual_seed(0)
        np.random.seed(0)
        mod.test_envs_obs_modes(env_id, obs_mode)
    finally:
        for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_record_episode_and_replay_tool(reference, tmp_path):
    """The reference's OWN `RecordEpisode` (mani_skill/utils/wrappers/record.py) and replay tool (mani_skill/trajectory/replay_trajectory.py `main`), unmodified, on
    the backend: 4 sub-scenes, episodes of 6 steps flushed by the vector wrapper's partial auto-reset into `<name>.h5` + `<name>.json` (the `h5py` the
    reference imports is compat/site/h5py where the real one is missing), read back with this repo's `load_trajectories`, then replayed by the reference's
    tool on 2 parallel sub-scenes with `--use-env-states`: the same episodes come out, the first transition bit-compatible."""
    gym = reference
    import h5py
    from mani_skill.trajectory.replay_trajectory import main, parse_args
    from mani_skill.utils.wrappers import RecordEpisode
    from mani_skill.vector.wrappers.gymnasium import ManiSkillVectorEnv
    from maniskill_b200.trajectory import load_trajectories
    out = str(tmp_path)
    env = gym.make("PickCube-v1", obs_mode="state_dict", num_envs=4, max_episode_steps=6)
    env = RecordEpisode(env, output_dir=out, trajectory_name="t", save_trajectory=True, save_video=False, record_reward=True)
    env = ManiSkillVectorEnv(env, max_episode_steps=6)
    env.reset(seed=5)
    for _ in range(9):
        env.step(env.action_space.sample())
    env.close()
    meta, trajs = load_trajectories(os.path.join(out, "t"))
    assert meta["env_info"]["env_id"] == "PickCube-v1" and len(meta["episodes"]) == 8 and sorted(trajs) == [f"traj_{i}" for i in range(8)]
    t0 = trajs["traj_0"]
    assert t0["actions"].shape == (6, 8) and t0["actions"].dtype == np.float32 and t0["rewards"].shape == (6,) and t0["truncated"][-1]
    assert t0["env_states"]["actors"]["cube"].shape == (7, 13) and t0["env_states"]["articulations"]["panda"].shape == (7, 13 + 9 * 2)
    assert t0["obs"]["agent"]["qpos"].shape == (7, 9)
    assert trajs["traj_4"]["actions"].shape == (3, 8)      # the episodes cut by close()
    main(parse_args(args=["--traj-path", os.path.join(out, "t.h5"), "--save-traj", "--use-env-states", "--sim-backend", "physx_cuda", "--num-envs", "2", "--allow-failure"]))
    replayed = os.path.join(out, "t.state_dict.pd_joint_delta_pos.physx_cuda")
    meta2, trajs2 = load_trajectories(replayed)
    assert len(meta2["episodes"]) == 8
    with h5py.File(os.path.join(out, "t.h5"), "r") as f:   # the container through the h5py surface the reference uses
        assert "traj_0" in f and f["traj_0"]["actions"][:].shape == (6, 8)
    by_len = lambda tr: sorted((v["actions"].shape[0], float(np.abs(v["actions"]).sum())) for v in tr.values())
    assert by_len(trajs) == by_len(trajs2)                  # the same episodes (the tool renumbers them in flush order)
    for v2 in trajs2.values():
        v1 = next(v for v in trajs.values() if v["actions"].shape == v2["actions"].shape and np.array_equal(v["actions"], v2["actions"]))
        # the tool's parallel path restores state t (not t + 1) after step t (replay_trajectory.py:217-222), so only the first transition starts from the
        # recorded state: s_1 = step(s_0, a_0) must be reproduced exactly (set_state_dict + step is deterministic on this backend)
        np.testing.assert_allclose(v2["env_states"]["actors"]["cube"][:2], v1["env_states"]["actors"]["cube"][:2], atol=1e-6)
        np.testing.assert_allclose(v2["env_states"]["articulations"]["panda"][:2], v1["env_states"]["articulations"]["panda"][:2], atol=1e-6)


def test_reference_own_test_partial_resets(reference):
    """/root/reference/tests/test_gpu_envs.py:245-270, executed as it is."""
    sys.modules.pop("tests", None)
    spec = importlib.util.spec_from_file_location("ref_test_gpu_envs", os.path.join(REF, "tests", "test_gpu_envs.py"))
    # the module does `from tests.utils import ...`: make `tests` resolve to the reference's tests package for the import
    ref_tests = importlib.util.spec_from_file_location("tests", os.path.join(REF, "tests", "__init__.py"), submodule_search_locations=[os.path.join(REF, "tests")])
    saved = {k: v for k, v in sys.modules.items() if k == "tests" or k.startswith("tests.")}
    try:
        pkg = importlib.util.module_from_spec(ref_tests)
        sys.modules["tests"] = pkg
        if os.path.exists(os.path.join(REF, "tests", "__init__.py")):
            ref_tests.loader.exec_module(pkg)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.test_partial_resets("PickCube-v1")
    finally:
        for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_own_test_raw_sim_states(reference):
    """/root/reference/tests/test_sim_state.py:10-40, executed as it is (state width 13 * 3 + 13 + 9 * 2 = 70)."""
    spec = importlib.util.spec_from_file_location("ref_test_sim_state", os.path.join(REF, "tests", "test_sim_state.py"))
    ref_tests = importlib.util.spec_from_file_location("tests", os.path.join(REF, "tests", "__init__.py"), submodule_search_locations=[os.path.join(REF, "tests")])
    saved = {k: v for k, v in sys.modules.items() if k == "tests" or k.startswith("tests.")}
    try:
        pkg = importlib.util.module_from_spec(ref_tests)
        sys.modules["tests"] = pkg
        if os.path.exists(os.path.join(REF, "tests", "__init__.py")):
            ref_tests.loader.exec_module(pkg)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.test_raw_sim_states()
    finally:
        for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_pick_cube_rollout_matches_the_mirror(reference):
    import maniskill_b200 as ms
    from emu_world import EmuBackendWorld
    gym = reference
    n = 4
    ref = gym.make("PickCube-v1", num_envs=n, obs_mode="state", sim_backend="physx_cuda")
    mir = ms.make("PickCube-v1", num_envs=n, obs_mode="state", world_factory=EmuBackendWorld)
    o1, _ = ref.reset(seed=5)
    o2, _ = mir.reset(seed=5)
    assert o1.shape == (n, 42) and float((o1 - o2).abs().max()) < 1e-5
    assert ref.unwrapped.get_state().shape == (n, 70)
    g = torch.Generator().manual_seed(0)
    for i in range(25):
        a = 2 * torch.rand((n, 8), generator=g) - 1
        o1, r1, te1, tr1, i1 = ref.step(a)
        o2, r2, te2, tr2, i2 = mir.step(a)
        assert float((o1 - o2).abs().max()) < 1e-4, i
        assert float((r1 - r2).abs().max()) < 1e-5, i
        assert torch.equal(i1["is_grasped"], i2["is_grasped"]) and torch.equal(i1["success"], i2["success"])
    ref.close()


@pytest.mark.parametrize("obs_mode", ["rgb+depth+segmentation", "state_dict"])
def test_reference_obs_modes(reference, obs_mode):
    gym = reference
    env = gym.make("PickCube-v1", num_envs=2, obs_mode=obs_mode, sim_backend="physx_cuda")
    obs, _ = env.reset(seed=0)
    obs, r, te, tr, info = env.step(torch.as_tensor(env.action_space.sample()))
    if obs_mode == "state_dict":
        assert obs["agent"]["qpos"].shape == (2, 9) and obs["extra"]["tcp_pose"].shape == (2, 7)
    else:
        sd = obs["sensor_data"]["base_camera"]
        assert sd["rgb"].shape == (2, 128, 128, 3) and sd["rgb"].dtype == torch.uint8
        assert sd["depth"].shape == (2, 128, 128, 1) and sd["depth"].dtype == torch.int16
        assert sd["segmentation"].shape == (2, 128, 128, 1) and sd["segmentation"].dtype == torch.int16
        ids = set(torch.unique(sd["segmentation"]).tolist())
        seg_map = env.unwrapped.segmentation_id_map
        names = {seg_map[i].name for i in ids if i in seg_map}
        assert "cube" in names and "table-workspace" in names and any(n.startswith("panda_link") for n in names), names
        assert sd["depth"].max() > 0
    env.close()


# every task of the reference whose assets ship inside the repository and whose scene this backend can express (fixed-base articulations,
# primitive / convex shapes): tabletop family, two-robot tasks, the dexterous-hand and valve tasks, other arms (SO100), an MJCF-built
# control task.  Not in the list: tasks that download assets (YCB, PartNet-Mobility, ReplicaCAD, Anymal / Unitree robots), free-floating
# articulation roots, the drawing tasks (hundreds of kinematic dots per sub-scene exceed the compiled capacities).  A scan of all 74 registered ids
# (num_envs=2, reset + one step) loads 36: the ones below, the further levels of RotateValve / TriFingerRotateCube (TriFinger's arena collides as its convex
# hull, with a warning) -- 24 need downloads, 6 have free-floating roots, 5 exceed capacities, 3 need packages that are not installed.
REFERENCE_TASKS = ["PushCube-v1", "StackCube-v1", "PullCube-v1", "LiftPegUpright-v1", "PokeCube-v1", "RollBall-v1", "PlaceSphere-v1", "StackPyramid-v1",
                   "PullCubeTool-v1", "PlugCharger-v1", "PegInsertionSide-v1", "PushT-v1", "TwoRobotPickCube-v1", "Empty-v1", "RotateValveLevel0-v1",
                   "RotateSingleObjectInHandLevel0-v1", "PickCubeSO100-v1", "SO100GraspCube-v1", "MS-CartpoleBalance-v1", "MS-CartpoleSwingUp-v1",
                   "MS-HopperHop-v1", "TwoRobotStackCube-v1", "OpenCabinetDrawer-v1", "OpenCabinetDoor-v1", "MS-HopperStand-v1", "RotateValveLevel1-v1",
                   "RotateSingleObjectInHandLevel1-v1", "TriFingerRotateCubeLevel0-v1", "TriFingerRotateCubeLevel4-v1"]


@pytest.mark.parametrize("task", REFERENCE_TASKS)
def test_reference_task_builds_and_steps(reference, task):
    """`gym.make(task, num_envs=2)` of the reference's own task module: its builders, URDF / MJCF loaders, agents and controllers record into the
    shim, `gpu_init()` compiles one batched world, reset + two control steps give finite observations and rewards without capacity overflow."""
    gym = reference
    env = gym.make(task, num_envs=2, obs_mode="state", sim_backend="physx_cuda")
    obs, _ = env.reset(seed=0)
    for _ in range(2):
        a = env.action_space.sample()
        a = {k: torch.as_tensor(v) for k, v in a.items()} if isinstance(a, dict) else torch.as_tensor(a)
        obs, r, te, tr, info = env.step(a)
    assert isinstance(obs, torch.Tensor) and obs.shape[0] == 2 and torch.isfinite(obs).all() and torch.isfinite(r).all()
    assert int(env.unwrapped.scene.px._world.overflow_flag.item()) == 0
    env.close()


def test_reference_open_cabinet_drawer(reference):
    """BASELINE.json configs[3]'s task through the reference's own module (mani_skill/envs/tasks/mobile_manipulation/open_cabinet_drawer.py), unchanged: the Fetch
    URDF (COLLADA visuals) through the reference's loader subclass, one PartNet-style cabinet per sub-scene (`set_scene_idxs([i])`, different model ids) merged
    with `Articulation.merge` / `Link.merge`, the handle position from the visual named `handle_*` (`Link.generate_mesh` + trimesh `center_mass`), the per-step
    `gpu_update_articulation_kinematics` / goal-site update.  The cabinets are stand-ins of the absent dataset (tools/make_standin_partnet.py).  Checked: shapes,
    closed drawers after reset, the goal site sits on the target handle, `open_enough` flips when the target drawer is pulled out."""
    gym = reference
    n = 4
    env = gym.make("OpenCabinetDrawer-v1", num_envs=n, obs_mode="state", sim_backend="physx_cuda")
    obs, _ = env.reset(seed=0)
    e = env.unwrapped
    assert obs.shape == (n, 44) and e.agent.robot.max_dof == 15 and e.cabinet.max_dof == 3 and len({c.name for c in e._cabinets}) == n
    assert e.get_state().shape == (n, 13 + 13 + 2 * 3 + 13 + 15 * 2) and set(e.handle_link.joint.type) == {"prismatic"}
    ql = e.cabinet.get_qlimits()
    assert torch.allclose(e.cabinet.qpos, ql[..., 0], atol=2e-3)
    assert float((e.handle_link_goal.pose.p - e.handle_link_positions()).abs().max()) < 1e-5
    assert float((e.handle_link_positions()[:, 0] - (-0.28)).abs().max()) < 5e-3          # the handle bar in front of the closed drawer (-D/2 - 0.03)
    for _ in range(2):
        obs, r, te, tr, info = env.step(torch.as_tensor(env.action_space.sample()))
    assert torch.isfinite(obs).all() and torch.isfinite(r).all() and not info["open_enough"].any()
    e.cabinet.set_qpos(ql[..., 1])
    e.scene._gpu_apply_all()
    e.scene.px.gpu_update_articulation_kinematics()
    e.scene._gpu_fetch_all()
    ev = e.evaluate()
    assert ev["open_enough"].all() and float((ev["handle_link_pos"][:, 0] - (-0.28 - 0.35)).abs().max()) < 5e-3
    assert int(e.scene.px._world.overflow_flag.item()) == 0
    env.close()


def test_reference_open_cabinet_door(reference):
    """OpenCabinetDoor-v1: the same module's subclass that targets revolute handles (every stand-in cabinet carries one door above its two drawers)."""
    gym = reference
    n = 4
    env = gym.make("OpenCabinetDoor-v1", num_envs=n, obs_mode="state", sim_backend="physx_cuda")
    env.reset(seed=0)
    e = env.unwrapped
    assert set(e.handle_link.joint.type) == {"revolute"}
    closed = e.handle_link_positions().clone()
    e.cabinet.set_qpos(e.cabinet.get_qlimits()[..., 1])
    e.scene._gpu_apply_all()
    e.scene.px.gpu_update_articulation_kinematics()
    e.scene._gpu_fetch_all()
    ev = e.evaluate()
    # the handle sits 0.69 m from the hinge: a quarter turn about the vertical hinge moves it out (-x) and across (+y), the height stays
    assert ev["open_enough"].all() and float((ev["handle_link_pos"][:, 2] - closed[:, 2]).abs().max()) < 1e-4
    assert float((ev["handle_link_pos"][:, 0] - closed[:, 0]).max()) < -0.5 and float((ev["handle_link_pos"][:, 1] - closed[:, 1]).min()) > 0.5
    env.close()


def test_reference_benchmark_script(reference, capsys):
    """mani_skill/examples/benchmarking/gpu_sim.py -- the script that defines BASELINE.json's metric (reset(seed=2022), 1000 steps of uniform actions, then 1000 steps
    with a reset every 200) -- `main(Args(...))` unmodified on the backend (examples/run_reference_benchmark.py is the launcher for a B200)."""
    from mani_skill.examples.benchmarking.gpu_sim import Args, main
    main(Args(env_id="PickCube-v1", obs_mode="state", num_envs=4, sim_freq=100, control_freq=20))
    out = capsys.readouterr().out
    assert "env.step:" in out and "env.step+env.reset:" in out and "4 parallel environments, sim_backend=physx_cuda" in out


def test_reference_reconfigure_rebuilds_the_world(reference):
    """`reset(options={"reconfigure": True})` (sapien_env.py:725-760; what tests/test_gpu_envs.py::test_env_reconfiguration exercises with YCB assets): the reference
    clears its scene and builds a new one -- a new batched world with newly drawn per-sub-scene peg geometry, which then steps."""
    gym = reference
    env = gym.make("PegInsertionSide-v1", num_envs=4, obs_mode="state", sim_backend="physx_cuda")
    env.reset(seed=1)
    e = env.unwrapped
    w1, h1 = e.scene.px._world, e.peg_half_sizes.clone()
    env.reset(seed=5, options=dict(reconfigure=True))
    assert e.scene.px._world is not w1 and float((h1 - e.peg_half_sizes).abs().max()) > 1e-3
    for _ in range(2):
        obs, r, _, _, _ = env.step(torch.as_tensor(env.action_space.sample()))
    assert obs.shape == (4, 43) and torch.isfinite(obs).all() and int(e.scene.px._world.overflow_flag.item()) == 0
    env.close()


def test_reference_wrappers_and_reward_modes(reference):
    """The reference's own `FlattenRGBDObservationWrapper` (mani_skill/utils/wrappers/flatten.py -- what its PPO-RGB baselines wrap), `ManiSkillVectorEnv` with
    `record_metrics` over the time limit (final_info / episode statistics at step 50, auto-reset), and the four reward modes, on the backend."""
    gym = reference
    from mani_skill.utils.wrappers.flatten import FlattenRGBDObservationWrapper
    from mani_skill.vector.wrappers.gymnasium import ManiSkillVectorEnv
    rewards = {}
    for mode in ("dense", "normalized_dense", "sparse", "none"):
        env = gym.make("PickCube-v1", num_envs=4, obs_mode="state", reward_mode=mode)
        env.reset(seed=0)
        rewards[mode] = env.step(torch.zeros(4, 8))[1].float()
        env.close()
    assert torch.allclose(rewards["normalized_dense"], rewards["dense"] / 5.0, atol=1e-6) and (rewards["sparse"] == 0).all() and (rewards["none"] == 0).all()
    env = FlattenRGBDObservationWrapper(gym.make("PickCube-v1", num_envs=4, obs_mode="rgbd"), rgb=True, depth=True, state=True)
    env = ManiSkillVectorEnv(env, auto_reset=True, ignore_terminations=True, record_metrics=True)
    obs, _ = env.reset(seed=0)
    assert obs["state"].shape == (4, 29) and obs["rgb"].shape == (4, 128, 128, 3) and obs["rgb"].dtype == torch.uint8 and obs["depth"].shape == (4, 128, 128, 1)
    for i in range(50):
        obs, r, te, tr, info = env.step(torch.as_tensor(env.action_space.sample()))
        assert ("final_info" in info) == (i == 49)
    assert tr.all() and not te.any() and (info["final_info"]["episode"]["episode_len"] == 50).all() and info["final_info"]["episode"]["return"].shape == (4,)
    assert (env.base_env.elapsed_steps == 0).all() and obs["rgb"].shape == (4, 128, 128, 3)     # already the first observation of the next episode
    env.close()
    # `env.render()` of the three render modes (sapien_env.py:1373-1440): the human camera, the tiled sensor pictures, both
    for mode, shape in (("rgb_array", (2, 512, 512, 3)), ("sensors", (2, 128, 256, 3)), ("all", (2, 512, 640, 3))):
        env = gym.make("PickCube-v1", num_envs=2, obs_mode="rgbd", render_mode=mode)
        env.reset(seed=0)
        img = env.render()
        assert tuple(img.shape) == shape and img.dtype == torch.uint8 and float(img.float().std()) > 5
        env.close()


@pytest.mark.parametrize("env_id", ["FrankaPickCubeBenchmark-v1", "FrankaMoveBenchmark-v1", "CartpoleBalanceBenchmark-v1"])
def test_reference_benchmark_environments(reference, env_id):
    """mani_skill/examples/benchmarking/envs: the environments the reference publishes its simulator comparisons with (gpu_sim.py's BENCHMARK_ENVS, configurable
    camera count / resolution) build, reset, step and render unchanged."""
    gym = reference
    import mani_skill.examples.benchmarking.envs  # noqa: F401
    env = gym.make(env_id, num_envs=4, obs_mode="rgb", num_cameras=2, camera_width=64, camera_height=48)
    obs, _ = env.reset(seed=0)
    obs, r, _, _, _ = env.step(torch.as_tensor(env.action_space.sample()))
    cams = obs["sensor_data"]
    assert len(cams) == 2 and all(c["rgb"].shape == (4, 48, 64, 3) and c["rgb"].dtype == torch.uint8 for c in cams.values())
    assert torch.isfinite(r).all() and int(env.unwrapped.scene.px._world.overflow_flag.item()) == 0
    env.close()


def test_reference_reset_to_env_states(reference):
    """`reset(options={"reset_to_env_states": {"env_states": ...}})` (sapien_env.py:935-944; SURVEY 8(c): how a cross-backend parity run copies states instead of
    relying on seeds), full and partial, with a state dictionary and with the flat state."""
    gym = reference
    env = gym.make("PickCube-v1", num_envs=4, obs_mode="state", sim_backend="physx_cuda")
    env.reset(seed=0)
    for _ in range(3):
        env.step(torch.as_tensor(env.action_space.sample()))
    e = env.unwrapped
    sd, flat, obs_then = e.get_state_dict(), e.get_state().clone(), e.get_obs().clone()
    for _ in range(3):
        env.step(torch.as_tensor(env.action_space.sample()))
    obs, _ = env.reset(options=dict(reset_to_env_states=dict(env_states=sd)))
    assert float((e.get_state() - flat).abs().max()) < 1e-6 and float((obs - obs_then).abs().max()) < 1e-4 and (e.elapsed_steps == 0).all()
    env.step(torch.as_tensor(env.action_space.sample()))
    moved = e.get_state().clone()
    idx = torch.tensor([1, 2])
    env.reset(options=dict(env_idx=idx, reset_to_env_states=dict(env_states=flat[idx])))
    now = e.get_state()
    assert float((now[idx] - flat[idx]).abs().max()) < 1e-6 and float((now[[0, 3]] - moved[[0, 3]]).abs().max()) < 1e-6
    env.close()


def test_reference_peg_insertion_rollout_matches_the_mirror(reference):
    """BASELINE.json configs[2]'s task: this repo's mirror (the path bench.py times for that side key) against the reference's own module on the same backend, same
    seed -- per-sub-scene peg / hole geometry from the same `_batched_episode_rng` draws, observations and rewards equal over a rollout."""
    import maniskill_b200 as ms
    from emu_world import EmuBackendWorld
    gym = reference
    n = 4
    ref = gym.make("PegInsertionSide-v1", num_envs=n, obs_mode="state", sim_backend="physx_cuda")
    mir = ms.make("PegInsertionSide-v1", num_envs=n, obs_mode="state", world_factory=EmuBackendWorld)
    o1, _ = ref.reset(seed=5)
    o2, _ = mir.reset(seed=5)
    assert o1.shape == (n, 43) and float((o1 - o2).abs().max()) < 1e-5
    assert float((ref.unwrapped.peg_half_sizes - torch.as_tensor(mir.peg_half_sizes)).abs().max()) < 1e-7
    g = torch.Generator().manual_seed(0)
    for i in range(12):
        a = 2 * torch.rand((n, 8), generator=g) - 1
        o1, r1, _, _, i1 = ref.step(a)
        o2, r2, _, _, i2 = mir.step(a)
        assert float((o1 - o2).abs().max()) < 1e-4 and float((r1 - r2).abs().max()) < 1e-5, i
        assert torch.equal(i1["success"], i2["success"])
    ref.close()
