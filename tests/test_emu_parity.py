"""CPU: the device per-env code compiled for the host (tests/emu) against the oracle -- proves the kernel logic
is the restated algorithm before any GPU time is spent. Both are float32 with contraction off => bit-level agreement."""
import numpy as np
import pytest

from scenarios import run_pick_cube


@pytest.mark.parametrize("emu_mode", [0, 1], ids=["fused", "pipelined"])
def test_emu_matches_oracle_f32_pick_cube(emu_mode):
    from maniskill_b200.scenes import pick_cube_scene
    cm = pick_cube_scene(6).compile()
    ref = run_pick_cube("oracle32", cm, 100)
    got = run_pick_cube("emu", cm, 100, emu_mode=emu_mode)
    assert np.abs(got["qpos"] - ref["qpos"]).max() < 1e-6
    assert np.abs(got["body"] - ref["body"]).max() < 1e-5
    rows = cm.link_rows["panda"]
    cube = cm.actor_rows["cube"]
    table = cm.actor_rows["table-workspace"]
    assert np.abs(got["world"].pair_impulse(cube, table) - ref["world"].pair_impulse(cube, table)).max() < 1e-6


@pytest.mark.parametrize("task,steps", [("PegInsertionSide-v1", 6), ("OpenCabinetDrawer-v1", 4)])
def test_emu_other_tasks_match_oracle(task, steps):
    """The pipelined substep on the scenes that use the other instantiations (per-env geometry overrides; two articulations,
    17 dofs, the 32-slot row records), device code compiled for the host against the oracle -- the CPU twin of
    tests/test_gpu_env.py::test_other_tasks_match_oracle."""
    from emu_world import EmuBackendWorld
    from scenarios import run_task_vs_oracle
    err_q, err_p, overflow = run_task_vs_oracle(task, steps, 3, world_factory=EmuBackendWorld)
    assert err_q < 1e-5 and err_p < 1e-5, (err_q, err_p)
    assert overflow == 0


def _edge_scenes():
    """Degenerate model shapes the pipeline has to take without special cases: no articulation at all (0 dofs), an articulation
    and nothing to collide with (0 candidate pairs, 0 bodies), bodies that only meet each other (no static geometry); a 48-vertex hull (cooked
    cylinder) settling on a box next to a falling box (GJK/EPA + hull-vertex patch, many candidates for reduce4)."""
    from test_oracle_kat import ground, link, root_link
    from maniskill_b200.model import SHAPE_BOX, SHAPE_SPHERE, ActorRec, ArticulationRec, SceneDesc, ShapeRec, SimParams, pose7
    n = 3
    s1 = SceneDesc(n, SimParams())
    ground(s1)
    s1.add_actor(ActorRec("box", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([0.05, 0.04, 0.03]))], pose7([0, 0, 0.2], [0.9238795, 0.3826834, 0, 0])))
    s1.add_actor(ActorRec("ball", "dynamic", [ShapeRec(SHAPE_SPHERE, pose7(), np.array([0.04, 0, 0]))], pose7([0.01, 0.0, 0.5])))
    robot = dict(name="arm", links=[root_link(), link("l1", 0, "revolute", (0, 0, 1.0), (0, 1, 0), 1.0, com=(0.2, 0, 0)),
                                    link("l2", 1, "prismatic", (0.4, 0, 0), (1, 0, 0), 0.5, lower=-0.05, upper=0.05)], disabled_collision_pairs=[])
    s2 = SceneDesc(n, SimParams())
    s2.add_articulation(ArticulationRec("arm", robot, pose7(), drive={"l1_joint": (50.0, 5.0, 1e10)}))
    s3 = SceneDesc(n, SimParams(gravity=(0, 0, 0)))
    for i, x in enumerate((-0.2, 0.2)):
        s3.add_actor(ActorRec(f"ball{i}", "dynamic", [ShapeRec(SHAPE_SPHERE, pose7(), np.array([0.05, 0, 0]))], pose7([x, 0.01 * i, 0]), angular_damping=0.0))
    from maniskill_b200.model import cylinder_shape
    s4 = SceneDesc(n, SimParams())
    s4.add_actor(ActorRec("slab", "static", [ShapeRec(SHAPE_BOX, pose7(), np.array([1.0, 1.0, 0.1]))], pose7([0, 0, -0.1])))
    # settling (not tumbling) contacts: an impact of a fast-rotating prism is chaotic -- the solver treats open and closed gaps
    # differently, so 1e-7 differences between the two code paths flip rows -- and tells nothing about parity
    s4.add_actor(ActorRec("cyl", "dynamic", [cylinder_shape(0.03, 0.05)], pose7([0, 0, 0.0325], [0.9990482, 0, 0, 0.0436194])))
    s4.add_actor(ActorRec("hullbox", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([0.03, 0.03, 0.03]))], pose7([0.3, 0.01, 0.05])))
    # convex mesh resting on a convex mesh (support-face patch of both hulls): a small prism standing on a larger one
    s5 = SceneDesc(n, SimParams())
    ground(s5)
    up = [0.7071068, 0, -0.7071068, 0]   # cylinder axis (local x) -> world z
    s5.add_actor(ActorRec("base", "dynamic", [cylinder_shape(0.12, 0.05)], pose7([0, 0, 0.0501], up)))
    s5.add_actor(ActorRec("top", "dynamic", [cylinder_shape(0.05, 0.04)], pose7([0.01, 0.005, 0.1405], up)))
    return [("bodies-only", s1, 120), ("articulation-only", s2, 120), ("two-bodies-no-static", s3, 80), ("hull-on-box", s4, 150), ("hull-on-hull", s5, 150)]


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4], ids=["bodies-only", "articulation-only", "two-bodies-no-static", "hull-on-box", "hull-on-hull"])
def test_emu_edge_models_match_oracle(idx):
    from emu import EmuWorld
    from oracle.oracle import OracleWorld
    name, scene, n_sub = _edge_scenes()[idx]
    cm = scene.compile()
    o, e = OracleWorld(cm, "f32"), EmuWorld(cm)
    nl = cm.scalars["n_link"]
    if idx == 1:
        tq = np.tile(np.array([[0.4, 0.02]]), (cm.scalars["n_envs"], 1))
        o.set_joint("target_qpos", tq)
        e.target_qpos[:] = tq
        e.apply(32)
    if idx == 2:  # head-on approach
        b = o.get_bodies()
        b[:, 0, 7], b[:, 1, 7] = 1.0, -0.5
        o.set_bodies(b)
        e.rigid_body_data[:, nl:] = b
        e.apply()
    o.step(n_sub)
    e.step(n_sub)
    ref, got = o.rigid_body_data(), e.rigid_body_data.astype(np.float64)
    assert np.isfinite(got).all()
    assert np.abs(got[..., :7] - ref[..., :7]).max() < 2e-5, np.abs(got[..., :7] - ref[..., :7]).max()
    if cm.scalars["n_dof"]:
        assert np.abs(e.qpos.astype(np.float64) - o.get_joint("qpos")).max() < 2e-5
    assert e.overflow() == 0


class _EmuAsOracle:
    """The few OracleWorld calls the known-answer tests use, on top of the emulated device code (pipelined substep)."""

    def __init__(self, cm):
        from emu import EmuWorld
        self.cm, self.w = cm, EmuWorld(cm)
        self.nl = cm.scalars["n_link"]

    def get_bodies(self):
        self.w.fetch()
        return self.w.rigid_body_data[:, self.nl:].astype(np.float64).copy()

    def set_bodies(self, b):
        self.w.rigid_body_data[:, self.nl:] = b
        self.w.apply()

    def step(self, n):
        self.w.step(n, 0)

    def pair_impulse(self, a, b):
        return self.w.pair_impulse(a, b).astype(np.float64)

    def set_joint(self, name, arr):
        buf = getattr(self.w, name)
        a = np.asarray(arr, dtype=np.float32).reshape(buf.shape[0], -1)
        buf[:, :a.shape[1]] = a
        self.w.apply()

    def get_joint(self, name):
        self.w.fetch()
        return getattr(self.w, name).astype(np.float64).copy()


def test_device_code_reproduces_the_analytic_answers_directly(monkeypatch):
    """Known-answer tests of tests/test_oracle_kat.py run on the emulated DEVICE code instead of the oracle: the ball that ends up rolling at
    5/7 v0, the stack of three boxes carrying 3 m g, the cooked cylinder resting on a box, the spinning box, and the two articulated chains against their
    Lagrangian equations (double pendulum; spatial chain with a prismatic joint)."""
    import test_oracle_kat as K
    monkeypatch.setattr(K, "OracleWorld", lambda cm, precision="f64": _EmuAsOracle(cm))
    K.test_sliding_ball_ends_up_rolling_at_five_sevenths_of_its_speed("box")
    K.test_sliding_ball_ends_up_rolling_at_five_sevenths_of_its_speed("plane")
    K.test_stack_of_three_boxes_stays_put()
    K.test_cooked_cylinder_rests_on_a_facet_and_weighs_mg("box", True)
    K.test_torsional_friction_of_a_spinning_box()
    # the articulated-body recursion of the DEVICE code (float32) against the independently integrated equations of motion
    K.test_double_pendulum_follows_the_lagrangian_equations()
    K.test_spatial_chain_follows_its_numerical_lagrangian()
