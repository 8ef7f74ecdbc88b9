"""CPU, only where the reference checkout exists (skipped elsewhere): methods of the reference's OWN `ManiSkillScene`
(mani_skill/envs/scene.py: `_gpu_apply_all`, `_gpu_fetch_all`, `get_pairwise_contact_impulses / _forces`) executed unmodified with `self.px` =
the `PhysxGpuSystem` facade of this repo (maniskill_b200/physx_shim.py) over the emulated backend.  The file is loaded from
/root/reference with its imports stubbed (sapien & co. are not installed); nothing of it is copied.  What the reference code computes
through the facade must equal what the BaseEnv mirror computes on its own path."""
import ast
import importlib.util
import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import pytest
import torch

import maniskill_b200 as ms
from emu_world import EmuBackendWorld

SCENE_PY = "/root/reference/mani_skill/envs/scene.py"
pytestmark = pytest.mark.skipif(not os.path.exists(SCENE_PY), reason="needs the reference checkout")


def _load_reference_module(path, as_name=None):
    """Execute one reference source file with every `mani_skill.*` / `sapien.*` import replaced by a MagicMock module (`as_name`: also
    register the result under that module name, so that a later file imports the real thing instead of a mock)."""
    def ensure(name):
        if name not in sys.modules:
            m = MagicMock(name=name)
            m.__name__, m.__path__, m.__all__ = name, [], []
            sys.modules[name] = m

    for node in ast.walk(ast.parse(open(path).read())):
        names = [node.module] if isinstance(node, ast.ImportFrom) and node.module else [a.name for a in node.names] if isinstance(node, ast.Import) else []
        for name in names:
            if name.split(".")[0] in ("mani_skill", "sapien", "trimesh", "gymnasium", "transforms3d", "dacite", "lxml", "pytorch_kinematics", "matplotlib", "h5py", "imageio", "tqdm", "PIL"):
                parts = name.split(".")
                for i in range(1, len(parts) + 1):
                    ensure(".".join(parts[:i]))
    spec = importlib.util.spec_from_file_location(as_name or "_reference_" + os.path.basename(path)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod          # dataclasses look their module up while the class body runs
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def reference_module():
    saved = dict(sys.modules)
    try:
        yield _load_reference_module
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
        sys.modules.update(saved)


@pytest.fixture()
def reference_scene_class(reference_module):
    return reference_module(SCENE_PY).ManiSkillScene


def test_reference_scene_methods_run_on_the_px_facade(reference_scene_class):
    RS = reference_scene_class
    n = 3
    ours, theirs = [ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode="pd_joint_pos", world_factory=EmuBackendWorld, fused=False) for _ in range(2)]
    for e in (ours, theirs):
        e.reset(seed=4)
    px = theirs.scene.px
    ref_scene = RS.__new__(RS)
    ref_scene.__dict__.update(px=px, gpu_sim_enabled=True, non_static_actors=[object()], articulations={"panda": object()}, _needs_fetch=False,
                              pairwise_contact_queries={}, _pairwise_contact_query_unique_hashes={})
    class Obj:   # what the query code reads off an Actor / Link: `.name`, `._bodies` (one body per sub-scene), hashability
        def __init__(self, name):
            self.name, self._bodies = name, [[b for b in px.bodies[e] if b.name == name][0] for e in range(n)]
    by_name = Obj
    cube, lf, rf, table = by_name("cube"), by_name("panda_panda_leftfinger"), by_name("panda_panda_rightfinger"), by_name("table-workspace")
    g = torch.Generator().manual_seed(2)
    for t in range(6):
        a = 2 * torch.rand(n, 8, generator=g) - 1
        a[:, 7] = -1.0                                       # close the gripper
        ours.step(a)
        # the reference's control step (sapien_env.py:1110-1131) with the reference's own scene methods
        theirs.agent.set_action(a)
        theirs.scene._dirty = 0
        RS._gpu_apply_all(ref_scene)                         # all eight px.gpu_apply_* calls
        assert ref_scene._needs_fetch
        for _ in range(5):
            px.step()
        RS._gpu_fetch_all(ref_scene)                         # all eight px.gpu_fetch_* calls
        assert not ref_scene._needs_fetch
        with pytest.raises(AssertionError):
            ref_scene._needs_fetch = True
            RS._gpu_apply_all(ref_scene)                     # the reference's guard against apply-apply without a fetch
        ref_scene._needs_fetch = False
        assert torch.equal(px.cuda_rigid_body_data.torch(), ours.scene.world.rigid_body_data)
        assert torch.equal(px.cuda_articulation_qpos.torch(), ours.scene.world.qpos)
    # contact queries through the reference's caching logic (query built once per pair of names, then re-run)
    for a_, b_, oa, ob in ((cube, table, ours.cube, ours.table), (lf, cube, ours.agent.finger1_link, ours.cube), (rf, cube, ours.agent.finger2_link, ours.cube)):
        imp = RS.get_pairwise_contact_impulses(ref_scene, a_, b_)
        assert imp.shape == (n, 3) and torch.allclose(imp, ours.scene.get_pairwise_contact_impulses(oa, ob), atol=1e-7)
        frc = RS.get_pairwise_contact_forces(ref_scene, a_, b_)
        assert torch.allclose(frc, ours.scene.get_pairwise_contact_forces(oa, ob), atol=1e-5)
    assert set(ref_scene.pairwise_contact_queries) == {"cubetable-workspace", "panda_panda_leftfingercube", "panda_panda_rightfingercube"}
    assert (RS.get_pairwise_contact_impulses(ref_scene, cube, table)[:, 2] > 0).all()     # the table pushes the cube up


def test_reference_rigid_body_struct_reads_through_the_px_facade(reference_module):
    """mani_skill/utils/structs/base.py (`PhysxRigidBodyComponentStruct`): `_body_data_index` from `body.gpu_pose_index`, velocities sliced out
    of `px.cuda_rigid_body_data`, the net-contact query built by `px.gpu_create_contact_body_impulse_query(self._bodies)` -- the reference's
    own accessors, with our body handles and facade underneath."""
    base = reference_module("/root/reference/mani_skill/utils/structs/base.py")
    RB = base.PhysxRigidBodyComponentStruct
    n = 3
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state", world_factory=EmuBackendWorld, fused=False)
    env.reset(seed=9)
    px = env.scene.px
    lifted = env.cube.pose.raw_pose.clone()
    lifted[1, :3] = torch.tensor([0.3, 0.3, 0.3])             # the cube of sub-scene 1 falls (clear of the gripper), the others rest
    from maniskill_b200.structs import Pose
    env.cube.set_pose(Pose(lifted))
    env.scene._gpu_apply_all()
    for _ in range(2):
        env.step(torch.zeros(n, 8))
    cube = RB.__new__(RB)
    cube.__dict__.update(scene=SimpleNamespace(px=px, gpu_sim_enabled=True, timestep=px.timestep, device=torch.device("cpu")), _body_data_name="cuda_rigid_body_data",
                         _body_data_index_internal=None, _bodies=[[b for b in px.bodies[e] if b.name == "cube"][0] for e in range(n)])
    assert cube._body_data_index.tolist() == [e * env.scene.world.n_rows + env.cube.row for e in range(n)]
    assert torch.equal(cube.linear_velocity, env.cube.linear_velocity) and torch.equal(cube.angular_velocity, env.cube.angular_velocity)
    assert float(cube.linear_velocity[1, 2]) == pytest.approx(-9.81 * 0.1, rel=2e-2) and float(cube.linear_velocity[0, 2].abs()) < 1e-3
    imp = cube.get_net_contact_impulses()
    assert torch.allclose(imp, env.scene.get_net_contact_impulses(env.cube), atol=1e-7)
    assert torch.equal(imp[1], torch.zeros(3)) and (imp[[0, 2], 2] > 0).all()           # in free fall nothing touches the cube
    assert torch.allclose(cube.get_net_contact_forces(), imp / 0.01)


def test_reference_articulation_struct_reads_and_writes_through_the_px_facade(reference_module):
    """mani_skill/utils/structs/articulation.py: `_data_index` from `gpu_index`, the `qpos` property (read and masked write into
    `px.cuda_articulation_qpos`), `get_joint_target_indices` + `set_joint_drive_targets` (meshgrid write into
    `px.cuda_articulation_target_qpos`) -- the reference's accessors on our handles; the mirror's Articulation sees the same numbers."""
    base = reference_module("/root/reference/mani_skill/utils/structs/base.py", as_name="mani_skill.utils.structs.base")
    sys.modules["mani_skill.utils.structs"].BaseStruct = base.BaseStruct          # the real class to inherit from, not a mock
    art_mod = reference_module("/root/reference/mani_skill/utils/structs/articulation.py")
    art_mod.common = SimpleNamespace(to_tensor=lambda x, device=None: torch.as_tensor(x, device=device))
    art_mod.ArticulationJoint = type("ArticulationJoint", (), {})        # isinstance() target; joints are addressed by index here
    RA = art_mod.Articulation
    n = 3
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode="pd_joint_pos", world_factory=EmuBackendWorld, fused=False)
    env.reset(seed=1)
    px = env.scene.px
    scene = SimpleNamespace(px=px, gpu_sim_enabled=True, device=torch.device("cpu"), _reset_mask=torch.tensor([True, False, True]))
    robot = RA.__new__(RA)
    robot.__dict__.update(scene=scene, _objs=[px.articulations[e][0] for e in range(n)], _scene_idxs=torch.arange(n), _cached_joint_target_indices={})
    assert robot._data_index.tolist() == [0, 1, 2] and robot.max_dof == 9
    assert torch.equal(robot.qpos, env.agent.robot.get_qpos())
    # masked write (sub-scenes 0 and 2, as during a partial reset), then apply + kinematics + fetch through the facade
    new_q = robot.qpos[[0, 2]].clone()
    new_q[:, 0] += 0.25
    before = env.agent.robot.get_qpos().clone()
    robot.qpos = new_q
    px.gpu_apply_articulation_qpos()
    px.gpu_update_articulation_kinematics()
    px.gpu_fetch_articulation_qpos()
    after = env.agent.robot.get_qpos()
    assert torch.allclose(after[[0, 2], 0], before[[0, 2], 0] + 0.25) and torch.equal(after[1], before[1])
    # drive targets of the arm joints of the masked sub-scenes
    arm = torch.arange(7, dtype=torch.int32)      # the controllers hold int32 joint indices (base_controller.py)
    gx, gy = robot.get_joint_target_indices(arm)
    assert gx.shape == (n, 7) and gy[0].tolist() == list(range(7))
    q_now = robot.qpos.clone()
    tgt = q_now[[0, 2], :7].clone()
    tgt[:, 0] += 0.3                                          # turn the first joint, hold the others
    robot.set_joint_drive_targets(tgt, joint_indices=arm)
    px.gpu_apply_articulation_target_position()
    px.gpu_fetch_articulation_target_qpos()
    t = px.cuda_articulation_target_qpos.torch()
    assert torch.allclose(t[[0, 2], :7], tgt) and abs(float(t[1, 0] - q_now[1, 0])) < 0.05
    for _ in range(60):
        px.step()
    px.gpu_fetch_articulation_qpos()
    q = robot.qpos
    assert (q[[0, 2], 0] - tgt[:, 0]).abs().max() < 0.02 and abs(float(q[1, 0] - q_now[1, 0])) < 0.02      # only the driven arms turned


def test_reference_texture_transforms_on_our_render_targets(reference_module):
    """mani_skill/render/shaders.py `PREBUILT_SHADER_CONFIGS["minimal"].texture_transforms`: the reference's own slicing of the raw `Color` /
    `PositionSegmentation` targets, applied to the targets of our camera group, equals the `sensor_data` the mirror delivers."""
    shaders = reference_module("/root/reference/mani_skill/render/shaders.py")
    cfg = shaders.PREBUILT_SHADER_CONFIGS["minimal"]
    assert cfg.texture_names == {"Color": ["rgb"], "PositionSegmentation": ["position", "depth", "segmentation"]}
    env = ms.make("PegInsertionSide-v1", num_envs=2, obs_mode="sensor_data", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    group = env._sensors.group
    for i, cam in enumerate(env._sensors.cams):
        want = {}
        for name, transform in cfg.texture_transforms.items():
            want.update(transform(group.get_picture_cuda(name, i)))
        got = obs["sensor_data"][cam["uid"]]
        assert set(got) == set(want) == {"rgb", "position", "depth", "segmentation"}
        for k in want:
            assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), (cam["uid"], k)


def _install_sapien():
    """`import sapien` = the shim package (maniskill_b200/compat/site/sapien), also where an earlier test left MagicMock modules under that name"""
    import maniskill_b200.compat as compat
    compat.install()
    for k in [k for k in sys.modules if k == "sapien" or k.startswith("sapien.")]:
        f = getattr(sys.modules[k], "__file__", None)
        if not (isinstance(f, str) and f.startswith(compat.SITE)):
            del sys.modules[k]
    import sapien
    return sapien


def test_reference_pose_struct_over_the_sapien_shim():
    """`compat.install()` makes `import sapien` resolve to maniskill_b200/compat/site/sapien; the reference's REAL batched `Pose`
    (mani_skill/utils/structs/pose.py, loaded with only its non-sapien imports stubbed) then accepts the shim's `sapien.Pose` objects:
    `Pose.create(sapien.Pose)`, `.sp` round trip, products against our own batched Pose."""
    saved = dict(sys.modules)
    try:
        sapien = _install_sapien()
        import sapien.physx as physx
        from maniskill_b200 import building
        assert issubclass(sapien.Pose, building.Pose) and physx.PhysxGpuSystem.__name__ == "PhysxGpuSystem"
        physx.enable_gpu()
        assert physx.is_gpu_enabled() and sapien.Device("cuda:1").is_cuda() and sapien.Device("cuda:1").cuda_id == 1 and sapien.Device("cpu").is_cpu()
        with pytest.raises(RuntimeError, match="no CPU simulation"):
            physx.PhysxCpuSystem()   # not part of the product: fails loudly rather than pretending
        rot = _load_reference_module("/root/reference/mani_skill/utils/geometry/rotation_conversions.py", as_name="mani_skill.utils.geometry.rotation_conversions")
        common = MagicMock()
        common.to_tensor = lambda x, device=None: torch.as_tensor(x, device=device).float() if not isinstance(x, torch.Tensor) else x.to(device)
        sys.modules["mani_skill.utils"] = MagicMock(common=common)
        sys.modules["mani_skill.utils.common"] = common
        pose_mod = _load_reference_module("/root/reference/mani_skill/utils/structs/pose.py")
        pose_mod.common = common
        RP = pose_mod.Pose
        a = sapien.Pose(p=[0.1, -0.2, 0.3], q=[0.9238795, 0, 0.3826834, 0])
        ra = RP.create(a)
        assert ra.raw_pose.shape == (1, 7) and torch.allclose(ra.p[0], torch.tensor([0.1, -0.2, 0.3])) and torch.allclose(ra.q[0], torch.tensor(a.q))
        b = sapien.Pose(p=[-0.3, 0.0, 0.2], q=[0.7071068, 0.7071068, 0, 0])
        prod_ref = (ra * RP.create(b)).raw_pose[0]
        prod_shim = a * b
        assert torch.allclose(prod_ref[:3], torch.tensor(prod_shim.p), atol=1e-6)
        assert min((prod_ref[3:] - torch.tensor(prod_shim.q)).abs().max(), (prod_ref[3:] + torch.tensor(prod_shim.q)).abs().max()) < 1e-6
        from maniskill_b200.structs import Pose as OurPose
        ours = (OurPose.create_from_pq(torch.tensor(a.p)[None], torch.tensor(a.q)[None]) * OurPose.create_from_pq(torch.tensor(b.p)[None], torch.tensor(b.q)[None])).raw_pose[0]
        assert torch.allclose(ours, prod_ref, atol=1e-6)
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
        sys.modules.update(saved)


TASK_FILES = {
    "PickCube-v1": ("pick_cube.py", "PickCubeEnv"), "PegInsertionSide-v1": ("peg_insertion_side.py", "PegInsertionSideEnv"),
    "OpenCabinetDrawer-v1": ("../mobile_manipulation/open_cabinet_drawer.py", "OpenCabinetDrawerEnv"),
}


@pytest.mark.parametrize("task", sorted(TASK_FILES))
def test_reference_task_logic_on_our_live_env(reference_module, task):
    """The reference's own task class -- `evaluate`, `_get_obs_extra`, `compute_dense_reward` of mani_skill/envs/tasks/tabletop/<task>.py --
    called with `self` = OUR running env (same attribute names: actors, agent, tcp, obs_mode_struct, device ...) at every step of a random
    rollout on the emulated backend; each result must equal what the mirror's own methods return on the same state."""
    from maniskill_b200 import structs
    fname, cls_name = TASK_FILES[task]
    _install_sapien()          # `sapien.Pose(...)` inside the task code is the shim's Pose, not a mock
    rot = reference_module("/root/reference/mani_skill/utils/geometry/rotation_conversions.py", as_name="mani_skill.utils.geometry.rotation_conversions")
    for name, attrs in (("mani_skill.envs.sapien_env", dict(BaseEnv=object)),
                        ("mani_skill.utils.registration", dict(register_env=lambda *a, **k: (lambda cls: cls))),
                        ("mani_skill.utils.structs.pose", dict(Pose=structs.Pose)), ("mani_skill.utils.structs", dict(Pose=structs.Pose)),
                        ("mani_skill.utils.geometry", dict(rotation_conversions=rot))):
        m = MagicMock(name=name, **attrs)
        m.__name__, m.__path__, m.__all__ = name, [], []
        sys.modules[name] = m
    mod = reference_module(f"/root/reference/mani_skill/envs/tasks/tabletop/{fname}")
    Ref = getattr(mod, cls_name)
    kw = dict(reward_mode="sparse") if task in ("StackPyramid-v1", "PlugCharger-v1") else {}
    env = ms.make(task, num_envs=3, obs_mode="state", world_factory=EmuBackendWorld, **kw)
    env.reset(seed=3)
    g = torch.Generator().manual_seed(0)
    for t in range(8):
        a = 2 * torch.rand(3, env.action_dim, generator=g) - 1
        a[:, -1] = -1.0 if t >= 2 else 1.0                     # close the gripper after two steps: contact forces on the fingers
        obs, rew, te, tr, info = env.step(a)
        status0 = env.reached_status.clone() if hasattr(env, "reached_status") else None
        ours_info = env.evaluate()
        ref_info = Ref.evaluate(env)
        assert set(ref_info) == set(ours_info), (set(ref_info), set(ours_info))
        for k in ref_info:
            assert torch.allclose(torch.as_tensor(ref_info[k]).float(), torch.as_tensor(ours_info[k]).float(), atol=1e-6), (t, k)
        ref_extra, ours_extra = Ref._get_obs_extra(env, ref_info), env._get_obs_extra(ours_info)
        assert list(ref_extra) == list(ours_extra)               # same keys in the same order: the flattened state vector has the same layout
        for k in ref_extra:
            assert torch.allclose(ref_extra[k].float(), ours_extra[k].float(), atol=1e-6), (t, k)
        if hasattr(Ref, "compute_dense_reward") and "compute_dense_reward" in Ref.__dict__:
            ours_r = env.compute_dense_reward(obs, a, ours_info)
            status1 = env.reached_status.clone() if status0 is not None else None
            if status0 is not None:
                env.reached_status = status0.clone()
            ref_r = Ref.compute_dense_reward(env, obs, a, {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in ref_info.items()})
            assert torch.allclose(ref_r, ours_r, atol=2e-5), (t, ref_r, ours_r)
            if status0 is not None:
                assert torch.equal(env.reached_status, status1)


INIT_TASKS = ["PickCube-v1", "PegInsertionSide-v1"]


@pytest.mark.parametrize("task", INIT_TASKS)
def test_reference_episode_initialisation_on_our_env(reference_module, task):
    """`reset(seed)` with the episode initialisation done by the REFERENCE's code: the task's own `_initialize_episode` and the real
    `TableSceneBuilder.initialize` (mani_skill/utils/scene_builder/table/scene_builder.py:68-127), `random_quaternions` and
    `UniformPlacementSampler` run against our env object (actors' `set_pose`, `agent.reset`, the episode RNG, `sapien.Pose` from the shim).
    The resulting simulation state must equal the state after the mirror's own reset with the same seed: same random streams, same
    layout, same robot configuration."""
    from maniskill_b200 import structs, utils as U
    fname, cls_name = TASK_FILES[task]
    _install_sapien()
    rot = reference_module("/root/reference/mani_skill/utils/geometry/rotation_conversions.py", as_name="mani_skill.utils.geometry.rotation_conversions")
    stubs = dict([("mani_skill.envs.sapien_env", dict(BaseEnv=object)), ("mani_skill.utils.registration", dict(register_env=lambda *a, **k: (lambda cls: cls))),
                  ("mani_skill.utils.structs.pose", dict(Pose=structs.Pose)), ("mani_skill.utils.structs", dict(Pose=structs.Pose)),
                  ("mani_skill.utils.geometry", dict(rotation_conversions=rot)), ("transforms3d", dict()), ("transforms3d.euler", dict(euler2quat=U.euler2quat)),
                  ("mani_skill.utils.scene_builder", dict(SceneBuilder=object)),
                  ("mani_skill.utils.common", dict(to_tensor=lambda x, device=None: torch.as_tensor(x, device=device)))])
    for name, attrs in stubs.items():
        m = MagicMock(name=name, **attrs)
        m.__name__, m.__path__, m.__all__ = name, [], []
        sys.modules[name] = m
    sys.modules["mani_skill.utils"] = MagicMock(common=sys.modules["mani_skill.utils.common"])
    rpose = reference_module("/root/reference/mani_skill/envs/utils/randomization/pose.py")
    rsamp = reference_module("/root/reference/mani_skill/envs/utils/randomization/samplers.py")
    rsamp.common = sys.modules["mani_skill.utils.common"]
    rcommon = reference_module("/root/reference/mani_skill/envs/utils/randomization/common.py")
    rcommon.common = sys.modules["mani_skill.utils.common"]
    rnd = MagicMock(random_quaternions=rpose.random_quaternions, UniformPlacementSampler=rsamp.UniformPlacementSampler, uniform=rcommon.uniform)
    rnd.__name__, rnd.__path__, rnd.__all__ = "mani_skill.envs.utils.randomization", [], []
    sys.modules["mani_skill.envs.utils.randomization"] = rnd
    sys.modules["mani_skill.envs.utils"] = MagicMock(randomization=rnd)
    table_mod = reference_module("/root/reference/mani_skill/utils/scene_builder/table/scene_builder.py")
    mod = reference_module(f"/root/reference/mani_skill/envs/tasks/tabletop/{fname}")
    mod.randomization = rnd
    Ref = getattr(mod, cls_name)
    kw = dict(reward_mode="sparse") if task in ("StackPyramid-v1", "PlugCharger-v1") else {}
    ours, theirs = [ms.make(task, num_envs=4, obs_mode="state", world_factory=EmuBackendWorld, **kw) for _ in range(2)]
    builder = table_mod.TableSceneBuilder.__new__(table_mod.TableSceneBuilder)
    builder.env, builder.table, builder.robot_init_qpos_noise = theirs, theirs.table, getattr(theirs, "robot_init_qpos_noise", 0.02)
    theirs.table_scene = theirs.scene_builder = builder      # (PullCubeTool-v1 calls it scene_builder)
    theirs._initialize_episode = lambda env_idx, options: Ref._initialize_episode(theirs, env_idx, options)
    for seed, idx in ((11, None), (12, torch.tensor([1, 3]))):
        opts = dict() if idx is None else dict(env_idx=idx)
        o1, _ = ours.reset(seed=seed, options=dict(opts))
        o2, _ = theirs.reset(seed=seed, options=dict(opts))
        s1, s2 = ours.get_state(), theirs.get_state()
        assert torch.allclose(s1, s2, atol=1e-6), (task, seed, float((s1 - s2).abs().max()), (s1 - s2).abs().max(dim=0).values.nonzero().flatten().tolist())
        assert torch.allclose(o1, o2, atol=1e-5)


def test_reference_panda_grasp_checks_on_our_agent_during_a_scripted_grasp(reference_module):
    """mani_skill/agents/robots/panda/panda.py `is_grasping` / `is_static` (with the real `common.compute_angle_between`) called with `self` =
    our agent while a scripted pick closes the gripper on the cube and lifts it: the reference's verdict follows ours through approach,
    grasp and lift (contact forces come from the px-level queries of the emulated device code)."""
    _install_sapien()
    for name, attrs in (("mani_skill.agents.base_agent", dict(BaseAgent=object, Keyframe=lambda **k: None)),
                        ("mani_skill.agents.registration", dict(register_agent=lambda *a, **k: (lambda cls: cls)))):
        m = MagicMock(name=name, **attrs)
        m.__name__, m.__path__, m.__all__ = name, [], []
        sys.modules[name] = m
    common = reference_module("/root/reference/mani_skill/utils/common.py", as_name="mani_skill.utils.common")
    sys.modules["mani_skill.utils"] = MagicMock(common=common)
    panda_mod = reference_module("/root/reference/mani_skill/agents/robots/panda/panda.py")
    panda_mod.common = common
    RefPanda = panda_mod.Panda
    n = 2
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode="pd_ee_target_delta_pos", world_factory=EmuBackendWorld)
    env.reset(seed=4)
    ctrl = env.agent.controller.controllers["arm"]
    base = torch.tensor([-0.615, 0.0, 0.0])
    cube0 = env.cube.pose.p.clone()
    seen = []

    def go_to(target, steps, grip, max_step=0.03):
        for _ in range(steps):
            a = torch.zeros(n, 4)
            a[:, :3] = (target - ctrl._target_pose.p).clamp(-max_step, max_step) / 0.1
            a[:, 3] = grip
            env.step(a)
            ours_g, ref_g = env.agent.is_grasping(env.cube), RefPanda.is_grasping(env.agent, env.cube)
            assert torch.equal(ours_g, ref_g)
            assert torch.equal(env.agent.is_static(0.2), RefPanda.is_static(env.agent, 0.2))
            for angle in (20, 85):
                assert torch.equal(env.agent.is_grasping(env.cube, max_angle=angle), RefPanda.is_grasping(env.agent, env.cube, max_angle=angle))
            seen.append(bool(ref_g.all()))

    go_to(cube0 - base + torch.tensor([0.0, 0.0, 0.10]), 12, 1.0)
    go_to(cube0 - base, 12, 1.0)
    assert not any(seen)
    go_to(cube0 - base, 8, -1.0)
    go_to(cube0 - base + torch.tensor([0.0, 0.0, 0.10]), 14, -1.0, max_step=0.015)
    assert seen[-1] and (env.cube.pose.p[:, 2] - cube0[:, 2] > 0.08).all()       # grasped and lifted, by the reference's own check


def test_reference_base_env_reset_drives_our_env(reference_module):
    """`BaseEnv.reset` of mani_skill/envs/sapien_env.py:857-978 itself -- seeding, reset mask, velocity clearing, episode initialisation
    under the forked torch RNG, `scene._gpu_apply_all()`, `scene.px.gpu_update_articulation_kinematics()`, `scene._gpu_fetch_all()`,
    controller reset, first observation -- executed with `self` = our env.  After the same sequence of seeded, partial and unseeded resets
    (with steps in between) it must leave the state and return the observation of the mirror's own `reset`."""
    gym = MagicMock(Env=type("Env", (), {}))
    gym.__name__, gym.__path__, gym.__all__ = "gymnasium", [], []
    sys.modules["gymnasium"] = gym
    common = reference_module("/root/reference/mani_skill/utils/common.py", as_name="mani_skill.utils.common")
    brng = reference_module("/root/reference/mani_skill/envs/utils/randomization/batched_rng.py")
    brng.common = common
    se = reference_module("/root/reference/mani_skill/envs/sapien_env.py")
    se.common, se.BatchedRNG = common, brng.BatchedRNG
    RefBaseEnv = se.BaseEnv
    for task in ("PickCube-v1",):
        ours, theirs = [ms.make(task, num_envs=4, obs_mode="state", world_factory=EmuBackendWorld) for _ in range(2)]
        theirs._batched_rng_backend = "numpy:random_state"
        g = torch.Generator().manual_seed(1)
        script = [dict(seed=21), dict(options=dict(env_idx=torch.tensor([0, 2]))), dict(seed=[5, 6, 7, 8]), dict(), dict(seed=3, options=dict(env_idx=torch.tensor([1])))]
        for kw in script:
            a = 2 * torch.rand(4, ours.action_dim, generator=g) - 1
            for e in (ours, theirs):
                e.step(a)
            torch.manual_seed(99)            # unseeded resets draw from the global torch stream: give both the same one
            o1, i1 = ours.reset(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
            torch.manual_seed(99)
            o2, i2 = RefBaseEnv.reset(theirs, **{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
            assert torch.allclose(o1, o2, atol=1e-5), (task, kw, (o1 - o2).abs().max(dim=1).values, (o1 - o2).abs().max(dim=0).values.nonzero().flatten())
            assert torch.allclose(ours.get_state(), theirs.get_state(), atol=1e-6), (task, kw)
            assert torch.equal(ours.elapsed_steps, theirs.elapsed_steps) and i2["reconfigure"] is False
            assert np.array_equal(np.asarray(ours._episode_seed), np.asarray(theirs._episode_seed))


def test_reference_base_env_step_drives_our_env(reference_module):
    """`BaseEnv.step` and `_step_action` of mani_skill/envs/sapien_env.py:1042-1132 themselves, with `self` = our env: the action goes
    through our agent, `self.scene.px.gpu_apply_articulation_target_position()`, five `self.scene.step()`, `self.scene._gpu_fetch_all()`, then
    `get_info` / `get_obs(info, unflattened=True)` / `get_reward` / `_flatten_raw_obs` and the termination logic of the reference.  Every
    return value equals the mirror's own `step` (one fused launch instead of five) over a rollout, in state and in state+rgb mode."""
    gym = MagicMock(Env=type("Env", (), {}))
    gym.__name__, gym.__path__, gym.__all__ = "gymnasium", [], []
    sys.modules["gymnasium"] = gym
    common = reference_module("/root/reference/mani_skill/utils/common.py", as_name="mani_skill.utils.common")
    se = reference_module("/root/reference/mani_skill/envs/sapien_env.py")
    se.common = common
    se.MultiAgent = type("MultiAgent", (), {})
    RefBaseEnv = se.BaseEnv
    calls = []
    for task, mode, cm in (("PickCube-v1", "state", "pd_joint_delta_pos"), ("PickCube-v1", "state_dict", "pd_joint_vel"), ("PickCube-v1", "state+rgb", "pd_joint_pos")):
        ours, theirs = [ms.make(task, num_envs=3, obs_mode=mode, control_mode=cm, world_factory=EmuBackendWorld, fused=False,
                                sensor_configs=dict(base_camera=dict(width=16, height=16))) for _ in range(2)]
        for e in (ours, theirs):
            e.reset(seed=8)
        theirs._step_action = (lambda e: (lambda action: RefBaseEnv._step_action(e, action)))(theirs)      # the reference's, not the mirror's
        px = theirs.scene.px
        for name in ("gpu_apply_articulation_target_position", "gpu_apply_articulation_target_velocity"):
            orig = getattr(px, name)
            setattr(px, name, (lambda o, n_: (lambda: (calls.append(n_), o())[1]))(orig, name))
        g = torch.Generator().manual_seed(5)
        for t in range(6):
            a = 2 * torch.rand(3, ours.action_dim, generator=g) - 1
            o1, r1, te1, tr1, i1 = ours.step(a)
            o2, r2, te2, tr2, i2 = RefBaseEnv.step(theirs, a)

            def same(x, y, path=""):
                if isinstance(x, dict):
                    assert set(x) == set(y), path
                    for k in x:
                        same(x[k], y[k], path + "/" + k)
                else:
                    assert x.dtype == y.dtype and torch.allclose(x.float(), y.float(), atol=1e-6), (task, mode, t, path)
            same(o1, o2, "obs")
            same(i1, i2, "info")
            assert torch.allclose(r1, r2, atol=1e-6) and torch.equal(te1, te2) and torch.equal(tr1, tr2)
            assert torch.equal(ours.scene.world.rigid_body_data, theirs.scene.world.rigid_body_data)
    assert "gpu_apply_articulation_target_position" in calls and "gpu_apply_articulation_target_velocity" in calls


def test_reference_vector_wrapper_and_timelimit_around_our_env(reference_module):
    """The reference's `TimeLimitWrapper.step` (mani_skill/utils/registration.py:127-170) and `ManiSkillVectorEnv.step / reset`
    (mani_skill/vector/wrappers/gymnasium.py:96-176) wrapped around OUR env, against the mirror's own vector wrapper around a twin env:
    observations, rewards, terminations, truncations, `final_info` / `final_observation`, the auto-reset of finished sub-scenes and the
    episode metrics agree over a rollout that crosses the time limit."""
    class Wrapper:                                   # gymnasium.Wrapper: forwards to `.env`
        def __init__(self, env):
            self.env = env
        def reset(self, *a, **k):
            return self.env.reset(*a, **k)
        def step(self, a):
            return self.env.step(a)
    gym = MagicMock(Wrapper=Wrapper, Env=type("Env", (), {}))
    gym.__name__, gym.__path__, gym.__all__ = "gymnasium", [], []
    sys.modules["gymnasium"] = gym
    vec = MagicMock(VectorEnv=object)
    vec.__name__, vec.__path__, vec.__all__ = "gymnasium.vector", [], []
    sys.modules["gymnasium.vector"] = vec
    common = reference_module("/root/reference/mani_skill/utils/common.py", as_name="mani_skill.utils.common")
    sys.modules["mani_skill.utils"] = MagicMock(common=common)
    reg = reference_module("/root/reference/mani_skill/utils/registration.py")
    vw = reference_module("/root/reference/mani_skill/vector/wrappers/gymnasium.py")
    vw.common = common
    n, limit = 3, 7
    ours_env, their_env = [ms.make("PickCube-v1", num_envs=n, obs_mode="state", world_factory=EmuBackendWorld) for _ in range(2)]
    their_env.unwrapped = their_env
    tl = reg.TimeLimitWrapper.__new__(reg.TimeLimitWrapper)
    tl.env, tl._max_episode_steps = their_env, limit
    tl.unwrapped = their_env
    ref = vw.ManiSkillVectorEnv.__new__(vw.ManiSkillVectorEnv)
    ref.__dict__.update(_env=tl, num_envs=n, auto_reset=True, ignore_terminations=False, record_metrics=True,
                        success_once=torch.zeros(n, dtype=torch.bool), fail_once=torch.zeros(n, dtype=torch.bool), returns=torch.zeros(n))
    ours = ms.ManiSkillVectorEnv(ours_env, auto_reset=True, record_metrics=True, max_episode_steps=limit)
    torch.manual_seed(0)
    o1, _ = ours.reset(seed=3)
    torch.manual_seed(0)
    o2, _ = ref.reset(seed=3)
    assert torch.allclose(o1, o2, atol=1e-6)
    g = torch.Generator().manual_seed(2)
    finals = 0
    for t in range(16):
        a = 2 * torch.rand(n, 8, generator=g) - 1
        torch.manual_seed(100 + t)                  # the unseeded auto-resets draw from the global stream
        r1 = ours.step(a)
        torch.manual_seed(100 + t)
        r2 = ref.step(a)
        for x, y in zip(r1[:4], r2[:4]):
            assert x.dtype == y.dtype and torch.allclose(x.float(), y.float(), atol=1e-6), t
        i1, i2 = r1[4], r2[4]
        assert ("final_info" in i1) == ("final_info" in i2)
        ep1 = i1["final_info"]["episode"] if "final_info" in i1 else i1["episode"]
        ep2 = i2["final_info"]["episode"] if "final_info" in i2 else i2["episode"]
        assert set(ep1) == set(ep2)
        for k in ep1:
            assert torch.allclose(ep1[k].float(), ep2[k].float(), atol=1e-6), (t, k)
        if "final_info" in i1:
            finals += 1
            assert torch.allclose(i1["final_observation"], i2["final_observation"], atol=1e-6) and torch.equal(i1["_final_info"], i2["_final_info"])
    assert finals == 2 and r1[3].sum() == 0 and int(ours_env.elapsed_steps[0]) == 2


def test_reference_record_episode_around_our_env(reference_module, tmp_path):
    """The reference's `RecordEpisode.reset / step / flush_trajectory` (mani_skill/utils/wrappers/record.py:356-756, h5py replaced by a
    dict-backed stand-in) recording OUR env -- observations, actions, rewards, flags and `get_state_dict()` of a real rollout with a partial
    reset -- writes exactly the datasets the mirror's recorder writes for a twin env."""
    from maniskill_b200.trajectory import RecordEpisode, _flatten, load_trajectories

    class Group:
        def __init__(self, store, path):
            self.store, self.path = store, path
        def create_group(self, name, track_order=True):
            return Group(self.store, f"{self.path}/{name}" if self.path else name)
        def create_dataset(self, key, data=None, dtype=None, **kw):
            self.store[f"{self.path}/{key}"] = np.array(data, dtype=dtype)

    class Wrapper:
        def __init__(self, env):
            self.env = env
        def reset(self, *a, **k):
            return self.env.reset(*a, **k)
        def step(self, a):
            return self.env.step(a)
    gym = MagicMock(Wrapper=Wrapper, Env=type("Env", (), {}))
    gym.__name__, gym.__path__, gym.__all__ = "gymnasium", [], []
    sys.modules["gymnasium"] = gym
    common = reference_module("/root/reference/mani_skill/utils/common.py", as_name="mani_skill.utils.common")
    sys.modules["mani_skill.utils"] = MagicMock(common=common, sapien_utils=SimpleNamespace(is_state_dict_consistent=lambda sd: True))
    sys.modules["mani_skill.utils.io_utils"] = MagicMock(dump_json=lambda *a, **k: None)
    rec_mod = reference_module("/root/reference/mani_skill/utils/wrappers/record.py")
    rec_mod.common, rec_mod.dump_json = common, (lambda *a, **k: None)
    rec_mod.sapien_utils = SimpleNamespace(is_state_dict_consistent=lambda sd: True)
    n = 3
    ours_env, their_env = [ms.make("PickCube-v1", num_envs=n, obs_mode="state", world_factory=EmuBackendWorld) for _ in range(2)]
    for e in (ours_env, their_env):
        e.max_episode_steps = None                   # the reference's recorder sees no TimeLimit here: compare the raw flags
    their_env.unwrapped = their_env
    their_env.get_wrapper_attr = lambda name: SimpleNamespace(sample=lambda: np.zeros(8, dtype=np.float32))
    store = {}
    ref = rec_mod.RecordEpisode.__new__(rec_mod.RecordEpisode)
    ref.env = their_env
    ref.__dict__.update(_h5_file=Group(store, ""), _json_data=dict(episodes=[]), _json_path="x.json", _trajectory_buffer=None, save_on_reset=True,
                        save_trajectory=True, record_env_state=True, record_reward=True, _episode_id=-1, _elapsed_record_steps=0, _save_video=False,
                        save_video_trigger=None, cpu_wrapped_env=False, _already_warned_about_state_dict_inconsistency=False, last_reset_kwargs={})
    ours = RecordEpisode(ours_env, str(tmp_path))
    g = torch.Generator().manual_seed(9)
    for r in (ours, ref):
        r.reset(seed=4)
    for t in range(7):
        a = 2 * torch.rand(n, 8, generator=g) - 1
        for r in (ours, ref):
            r.step(a)
        if t == 3:
            torch.manual_seed(1)
            ours.reset(options=dict(env_idx=torch.tensor([2])))
            torch.manual_seed(1)
            ref.reset(options=dict(env_idx=torch.tensor([2])))
    ours.flush_trajectory()
    ours._dump()
    ref.flush_trajectory()
    meta, trajs = load_trajectories(str(tmp_path / "trajectory"))
    flat = {}
    for name, tr in trajs.items():
        _flatten(name, tr, flat)
    assert set(flat) == set(store) and len(store) == 4 * 10       # four episodes x (obs, actions, 3 flags, rewards, 4 state arrays)
    for k in store:
        assert flat[k].dtype == store[k].dtype and flat[k].shape == store[k].shape and np.allclose(flat[k], store[k], atol=1e-6), k
    assert [e["elapsed_steps"] for e in meta["episodes"]] == [e["elapsed_steps"] for e in ref._json_data["episodes"]] == [4, 7, 7, 3]
    assert [e["episode_seed"] for e in meta["episodes"]] == [int(e["episode_seed"]) for e in ref._json_data["episodes"]]


def test_reference_flatten_wrappers_on_our_live_observations(reference_module):
    """mani_skill/utils/wrappers/flatten.py: the reference's `FlattenRGBDObservationWrapper.observation` and `FlattenObservationWrapper.observation`
    applied to observations produced by OUR env (two cameras, visual and state_dict modes) equal the mirror's wrappers."""
    from maniskill_b200.wrappers import FlattenObservationWrapper, FlattenRGBDObservationWrapper

    class ObsWrapper:
        def __init__(self, env):
            self.env = env
    gym = MagicMock(ObservationWrapper=ObsWrapper, ActionWrapper=ObsWrapper, Env=type("Env", (), {}))
    gym.__name__, gym.__path__, gym.__all__ = "gymnasium", [], []
    sys.modules["gymnasium"] = gym
    common = reference_module("/root/reference/mani_skill/utils/common.py", as_name="mani_skill.utils.common")
    sys.modules["mani_skill.utils"] = MagicMock(common=common)
    fl = reference_module("/root/reference/mani_skill/utils/wrappers/flatten.py")
    fl.common = common
    for mode in ("rgbd", "state+rgb+depth"):
        env = ms.make("PegInsertionSide-v1", num_envs=2, obs_mode=mode, world_factory=EmuBackendWorld, sensor_configs=dict(base_camera=dict(width=16, height=16), hand_camera=dict(width=16, height=16)))
        obs, _ = env.reset(seed=0)
        for sep in (True, False):
            ref_self = SimpleNamespace(base_env=SimpleNamespace(device=torch.device("cpu")), include_rgb=True, include_depth=True, sep_depth=sep, include_state=True)
            want = fl.FlattenRGBDObservationWrapper.observation(ref_self, {k: (dict(v) if isinstance(v, dict) else v) for k, v in obs.items()})
            got = FlattenRGBDObservationWrapper(env, sep_depth=sep).observation(obs)
            assert set(got) == set(want)
            for k in want:
                assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), (mode, sep, k)
    sd = ms.make("PickCube-v1", num_envs=2, obs_mode="state_dict", world_factory=EmuBackendWorld)
    o, _ = sd.reset(seed=0)
    assert torch.equal(FlattenObservationWrapper(sd).observation(o), fl.FlattenObservationWrapper.observation(SimpleNamespace(), o))


def test_reference_fetch_checks_on_our_agent(reference_module):
    """mani_skill/agents/robots/fetch/fetch.py `is_static` / `is_grasping` with `self` = our Fetch while it drives around and moves its arm
    (OpenCabinetDrawer-v1): same verdicts as the mirror, including the base-velocity threshold."""
    _install_sapien()
    for name, attrs in (("mani_skill.agents.base_agent", dict(BaseAgent=object, Keyframe=lambda **k: None, DictControllerConfig=dict)),
                        ("mani_skill.agents.registration", dict(register_agent=lambda *a, **k: (lambda cls: cls)))):
        m = MagicMock(name=name, **attrs)
        m.__name__, m.__path__, m.__all__ = name, [], []
        sys.modules[name] = m
    common = reference_module("/root/reference/mani_skill/utils/common.py", as_name="mani_skill.utils.common")
    sys.modules["mani_skill.utils"] = MagicMock(common=common)
    fetch_mod = reference_module("/root/reference/mani_skill/agents/robots/fetch/fetch.py")
    fetch_mod.common = common
    RefFetch = fetch_mod.Fetch
    env = ms.make("OpenCabinetDrawer-v1", num_envs=2, obs_mode="state", world_factory=EmuBackendWorld)
    env.reset(seed=0)
    g = torch.Generator().manual_seed(0)
    verdicts = []
    for t in range(10):
        a = torch.zeros(2, env.action_dim) if t < 3 else 2 * torch.rand(2, env.action_dim, generator=g) - 1
        env.step(a)
        ours_s, ref_s = env.agent.is_static(0.2), RefFetch.is_static(env.agent, 0.2)
        assert torch.equal(ours_s, ref_s), t
        assert torch.equal(env.agent.is_grasping(env.handle_link), RefFetch.is_grasping(env.agent, env.handle_link))
        verdicts.append(bool(ref_s.all()))
    assert verdicts[1] and not all(verdicts)           # at rest first, moving later
