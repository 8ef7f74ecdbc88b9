"""CPU: known-answer tests that pin the oracle (PhysX is absent, so the physics restatement is
anchored on analytic results instead of golden vectors)."""
import numpy as np
import pytest

from maniskill_b200.model import (SHAPE_BOX, SHAPE_PLANE, SHAPE_SPHERE, ActorRec, ArticulationRec, SceneDesc, ShapeRec, SimParams, pose7)
from oracle.oracle import OracleWorld, collide

DT = 0.01
G = 9.81


def link(name, parent, jtype, p, axis, mass, com=(0, 0, 0), inertia=(1e-3, 1e-3, 1e-3, 0, 0, 0), lower=-1e30, upper=1e30):
    return dict(name=name, parent=parent, mass=mass, com=list(com), inertia=list(inertia), collisions=[],
                joint=dict(name=name + "_joint", type=jtype, p=list(p), q=[1, 0, 0, 0], axis=list(axis), lower=lower, upper=upper,
                           effort=0, damping=0, friction=0))


def root_link():
    return dict(name="base", parent=-1, mass=1.0, com=[0, 0, 0], inertia=[1, 1, 1, 0, 0, 0], collisions=[],
                joint=dict(name="", type="fixed", p=[0, 0, 0], q=[1, 0, 0, 0], axis=[1, 0, 0], lower=0, upper=0, effort=0, damping=0, friction=0))


def ground(scene, mu=0.3):
    scene.add_actor(ActorRec("ground", "static", [ShapeRec(SHAPE_PLANE, pose7([0, 0, 0], [0.7071068, 0, -0.7071068, 0]), mu=mu)], pose7()))


def test_free_fall_matches_substepped_euler():
    s = SceneDesc(1, SimParams())
    s.add_actor(ActorRec("ball", "dynamic", [ShapeRec(SHAPE_SPHERE, pose7(), np.array([0.05, 0, 0]))], pose7([0, 0, 10.0]), angular_damping=0.0))
    w = OracleWorld(s.compile(), "f64")
    n = 50
    w.step(n)
    b = w.get_bodies()[0, 0]
    assert b[9] == pytest.approx(-G * DT * n, rel=1e-6)  # model tables are float32 (dt, g)
    # positions integrate the velocity of each of the 15 sub-steps (h = dt/15): per step dx = dt v + g dt^2 (15+1)/(2*15)
    assert b[2] == pytest.approx(10.0 - G * DT * DT * (n * (n - 1) / 2 + n * 16 / 30), rel=1e-6)


def test_pd_drive_is_the_implicit_recurrence():
    m, kp, kd = 2.0, 1e3, 1e2
    robot = dict(name="slider", links=[root_link(), link("mass", 0, "prismatic", (0, 0, 0), (1, 0, 0), m)], disabled_collision_pairs=[])
    s = SceneDesc(1, SimParams())
    s.add_articulation(ArticulationRec("slider", robot, pose7(), drive={"mass_joint": (kp, kd, 1e10)}))
    w = OracleWorld(s.compile(), "f64")
    w.set_joint("target_qpos", [[0.3]])
    q, v = 0.0, 0.0
    for _ in range(60):
        w.step(1)
        v_new = (m * v + DT * (kp * (0.3 - q))) / (m + DT * kd + DT * DT * kp)  # implicit spring-damper, target velocity 0
        q += DT * v + (v_new - v) * DT * 16 / 30  # the acceleration is applied over 15 sub-steps
        v = v_new
        assert w.get_joint("qpos")[0, 0] == pytest.approx(q, abs=1e-6)
    assert abs(q - 0.3) < 0.02


def test_drive_force_limit_caps_acceleration():
    m, fl = 2.0, 5.0
    robot = dict(name="slider", links=[root_link(), link("mass", 0, "prismatic", (0, 0, 0), (1, 0, 0), m)], disabled_collision_pairs=[])
    s = SceneDesc(1, SimParams())
    s.add_articulation(ArticulationRec("slider", robot, pose7(), drive={"mass_joint": (1e3, 1e2, fl)}))
    w = OracleWorld(s.compile(), "f64")
    w.set_joint("target_qpos", [[10.0]])
    w.step(10)
    assert w.get_joint("qvel")[0, 0] == pytest.approx(fl / m * DT * 10, rel=1e-6)


def test_pendulum_period_and_energy():
    L, m = 0.5, 1.0
    I = 1e-4
    robot = dict(name="pend", links=[root_link(), link("bob", 0, "revolute", (0, 0, 0), (0, 1, 0), m, com=(0, 0, -L), inertia=(I, I, I, 0, 0, 0))],
                 disabled_collision_pairs=[])
    s = SceneDesc(1, SimParams(sim_freq=1000))
    s.add_articulation(ArticulationRec("pend", robot, pose7(), disable_gravity=False))
    w = OracleWorld(s.compile(), "f64")
    th0 = 0.1
    w.set_joint("qpos", [[th0]])
    qs = []
    for _ in range(3000):
        w.step(1)
        qs.append(w.get_joint("qpos")[0, 0])
    qs = np.array(qs)
    zc = np.where((qs[:-1] > 0) & (qs[1:] <= 0))[0]
    period = (zc[1] - zc[0]) * 1e-3
    analytic = 2 * np.pi * np.sqrt((I + m * L * L) / (m * G * L))
    assert period == pytest.approx(analytic, rel=5e-3)
    assert np.abs(qs).max() == pytest.approx(th0, rel=2e-2)  # semi-implicit Euler: bounded energy error


def test_box_rests_on_plane_and_weighs_mg():
    s = SceneDesc(1, SimParams())
    ground(s)
    s.add_actor(ActorRec("box", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([0.05, 0.05, 0.05]))], pose7([0, 0, 0.05])))
    cm = s.compile()
    w = OracleWorld(cm, "f64")
    w.step(200)
    b = w.get_bodies()[0, 0]
    assert abs(b[2] - 0.05) < 1e-3 and np.abs(b[:2]).max() < 1e-5 and np.abs(b[7:]).max() < 1e-2
    mass = 1000 * 0.1**3
    imp = w.pair_impulse(cm.actor_rows["box"], -1)
    assert imp[0, 2] == pytest.approx(mass * G * DT, rel=1e-3)


@pytest.mark.parametrize("angle_deg,slides", [(10.0, False), (25.0, True)])
def test_coulomb_threshold(angle_deg, slides):
    """mu = 0.3 on both -> combined 0.3: tan(16.7 deg) = 0.3.  Tilt gravity instead of the plane."""
    th = np.deg2rad(angle_deg)
    s = SceneDesc(1, SimParams(gravity=(G * np.sin(th), 0, -G * np.cos(th))))
    ground(s)
    s.add_actor(ActorRec("box", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([0.05, 0.05, 0.02]))], pose7([0, 0, 0.02])))
    w = OracleWorld(s.compile(), "f64")
    w.step(100)
    b = w.get_bodies()[0, 0]
    if slides:
        a = G * (np.sin(th) - 0.3 * np.cos(th))
        assert b[7] == pytest.approx(a * 1.0, rel=0.05)
    else:
        assert abs(b[0]) < 1e-3 and abs(b[7]) < 1e-3


def test_momentum_is_conserved_in_a_collision():
    s = SceneDesc(1, SimParams(gravity=(0, 0, 0)))
    for i, (x, vx) in enumerate([(-0.2, 1.0), (0.2, -0.5)]):
        s.add_actor(ActorRec(f"ball{i}", "dynamic", [ShapeRec(SHAPE_SPHERE, pose7(), np.array([0.05, 0, 0]), density=1000 * (1 + i))],
                             pose7([x, 0, 0]), angular_damping=0.0))
    cm = s.compile()
    w = OracleWorld(cm, "f64")
    b = w.get_bodies()
    b[0, 0, 7], b[0, 1, 7] = 1.0, -0.5
    w.set_bodies(b)
    masses = cm.arrays["fb_mass"]
    p0 = masses[0] * 1.0 + masses[1] * -0.5
    w.step(60)
    b = w.get_bodies()[0]
    assert masses[0] * b[0, 7] + masses[1] * b[1, 7] == pytest.approx(p0, rel=1e-6)
    assert b[0, 7] <= b[1, 7] + 1e-6  # they no longer approach (restitution 0)


def test_narrowphase_known_answers():
    box = dict(type=SHAPE_BOX, pose=[0, 0, 0.049, 1, 0, 0, 0], size=[0.05, 0.05, 0.05])
    plane = dict(type=SHAPE_PLANE, pose=[0, 0, 0, 0.7071068, 0, -0.7071068, 0])
    c = collide(box, plane)
    assert len(c) == 4 and np.allclose(c[:, 6], -0.001, atol=1e-7) and np.allclose(c[:, 3:6], [0, 0, 1], atol=1e-6)
    b2 = dict(type=SHAPE_BOX, pose=[0, 0, 0.1 + 0.002, 1, 0, 0, 0], size=[0.05, 0.05, 0.05])
    b1 = dict(type=SHAPE_BOX, pose=[0, 0, 0, 1, 0, 0, 0], size=[0.05, 0.05, 0.05])
    c = collide(b2, b1)
    assert len(c) == 4 and np.allclose(c[:, 6], 0.002, atol=1e-6) and np.allclose(c[:, 3:6], [0, 0, 1], atol=1e-6)
    sph_a = dict(type=SHAPE_SPHERE, pose=[0.0, 0, 0, 1, 0, 0, 0], size=[0.05, 0, 0])
    sph_b = dict(type=SHAPE_SPHERE, pose=[0.12, 0, 0, 1, 0, 0, 0], size=[0.05, 0, 0])
    c = collide(sph_a, sph_b)
    assert len(c) == 1 and c[0, 6] == pytest.approx(0.02, abs=1e-6) and np.allclose(c[0, 3:6], [-1, 0, 0], atol=1e-6)
    # GJK/EPA: an octahedron hull penetrating a box by 5 mm along z
    octa = np.array([[0.05, 0, 0], [-0.05, 0, 0], [0, 0.05, 0], [0, -0.05, 0], [0, 0, 0.05], [0, 0, -0.05]], dtype=np.float32)
    hull = dict(type=4, pose=[0, 0, 0.095, 1, 0, 0, 0], size=[0, 0, 0], verts=octa)
    c = collide(hull, b1)
    assert len(c) == 1 and c[0, 6] == pytest.approx(-0.005, abs=1e-5) and np.allclose(c[0, 3:6], [0, 0, 1], atol=1e-4)


@pytest.mark.parametrize("support,tilt", [("plane", False), ("box", False), ("box", True)])
def test_cooked_cylinder_rests_on_a_facet_and_weighs_mg(support, tilt):
    """add_cylinder_collision (actor_builder.py:104-116) = a cooked convex prism, axis local x.  A 48-vertex hull resting on the
    ground plane (vertex tests) or on a box (GJK/EPA + the hull-vertex patch) must stay put for hundreds of steps and load its
    support with exactly m g dt: the regression for the near-contact manifold selection (reduce4) and hull_box_patch."""
    from maniskill_b200.model import cylinder_shape
    r, hl = 0.03, 0.05
    shape = cylinder_shape(r, hl, density=1000.0)
    m, c, I = shape.mass_props()
    assert len(shape.vertices) == 48 and np.abs(c).max() < 1e-9
    assert m == pytest.approx(1000.0 * np.pi * r * r * 2 * hl, rel=0.02)
    assert I[0, 0] == pytest.approx(0.5 * m * r * r, rel=0.03) and I[1, 1] == pytest.approx(m * (r * r / 4 + (2 * hl) ** 2 / 12), rel=0.03)
    s = SceneDesc(1, SimParams())
    if support == "plane":
        ground(s)
    else:
        s.add_actor(ActorRec("slab", "static", [ShapeRec(SHAPE_BOX, pose7(), np.array([1.0, 1.0, 0.1]))], pose7([0, 0, -0.1])))
    q = [0.9659258, 0, 0.2588190, 0] if tilt else [1, 0, 0, 0]  # 30 degrees about y: lands on the rim of an end cap first
    s.add_actor(ActorRec("cyl", "dynamic", [shape], pose7([0, 0, r + 0.04], q)))
    cm = s.compile()
    w = OracleWorld(cm, "f64")
    w.step(600)
    b = w.get_bodies()[0, 0]
    assert abs(b[2] - r * np.cos(np.pi / 24)) < 1e-3, b[:3]        # lies on a facet of the prism
    assert np.abs(b[7:10]).max() < 5e-3 and np.abs(b[10:13]).max() < 5e-2
    assert w.pair_impulse(cm.actor_rows["cyl"], -1)[0, 2] == pytest.approx(m * G * DT, rel=1e-2)


def test_stack_of_three_boxes_stays_put():
    """Solver convergence on a contact chain: three 10 cm boxes stacked on the ground keep their heights for 400 steps and the ground
    carries the weight of all three."""
    s = SceneDesc(1, SimParams())
    ground(s)
    for i in range(3):
        s.add_actor(ActorRec(f"box{i}", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([0.05, 0.05, 0.05]))], pose7([0.002 * i, -0.001 * i, 0.05 + 0.1 * i])))
    cm = s.compile()
    w = OracleWorld(cm, "f64")
    w.step(400)
    b = w.get_bodies()[0]
    for i in range(3):
        # the soft contacts (30 Hz) sag by about a millimetre per loaded interface; no drift beyond that
        assert abs(b[i, 2] - (0.05 + 0.1 * i)) < 1.5e-3 * (i + 1) and np.abs(b[i, :2] - [0.002 * i, -0.001 * i]).max() < 2e-3, (i, b[i, :3])
        assert np.abs(b[i, 7:]).max() < 2e-2
    m = 1000 * 0.1 ** 3
    assert w.pair_impulse(cm.actor_rows["box0"], -1)[0, 2] == pytest.approx(3 * m * G * DT, rel=2e-2)
    assert w.pair_impulse(cm.actor_rows["box1"], cm.actor_rows["box0"])[0, 2] == pytest.approx(2 * m * G * DT, rel=2e-2)


def test_joint_limit_stops_a_falling_link():
    """A horizontal link released under gravity swings down until its joint limit (-0.5 rad) and rests there; the limit row only
    acts when it is reached."""
    robot = dict(name="arm", links=[root_link(), link("l1", 0, "revolute", (0, 0, 1.0), (0, 1, 0), 1.0, com=(0.3, 0, 0), lower=-0.5, upper=0.5)],
                 disabled_collision_pairs=[])
    s = SceneDesc(1, SimParams())
    s.add_articulation(ArticulationRec("arm", robot, pose7(), disable_gravity=False))
    w = OracleWorld(s.compile(), "f64")
    q_hist = []
    for _ in range(300):
        w.step(1)
        q_hist.append(w.get_joint("qpos")[0, 0])
    q_hist = np.array(q_hist)
    # rotation about +y with the centre of mass along +x: gravity drives q upwards (towards +0.5); free fall first, then the stop
    assert q_hist[5] > 0 and q_hist[5] == pytest.approx(0.5 * (G / 0.3) * (6 * DT) ** 2, rel=0.25)
    assert q_hist.max() < 0.5 + 5e-3
    assert abs(q_hist[-1] - 0.5) < 2e-3 and abs(w.get_joint("qvel")[0, 0]) < 1e-2


def test_mimic_tendon_keeps_the_fingers_mirrored():
    """The fixed-tendon coupling of the Panda gripper (`<mimic>` of finger_joint2, articulation_builder.py:161-200): with only one
    finger driven towards closing and an obstacle-free world, both finger joints move together."""
    from maniskill_b200.scenes import pick_cube_scene, PANDA_REST_QPOS
    cm = pick_cube_scene(1).compile()
    w = OracleWorld(cm, "f32")
    q0 = PANDA_REST_QPOS.copy()[None]
    w.set_joint("qpos", q0)
    tq = q0.copy()
    tq[0, 7] = 0.0   # only finger 1 is told to close; finger 2 keeps its open target
    w.set_joint("target_qpos", tq)
    w.step(100)
    q = w.get_joint("qpos")[0]
    assert abs(q[7] - q[8]) < 2e-3, q[7:]
    assert 0.005 < q[7] < 0.035  # the two drives (one closing, one holding open) balance through the tendon


@pytest.mark.parametrize("support", ["plane", "box"])
def test_sliding_ball_ends_up_rolling_at_five_sevenths_of_its_speed(support):
    """A solid sphere thrown along the ground without spin: Coulomb friction at the contact point slows the centre and spins the
    ball up until v = omega r, which for I = 2/5 m r^2 happens at v = 5/7 v0, and from then on nothing changes (no rolling
    resistance, like PhysX).  Checks the friction rows' lever arm / angular coupling and the sphere paths of the narrowphase."""
    r, v0 = 0.035, 1.0
    s = SceneDesc(1, SimParams())
    if support == "plane":
        ground(s, mu=0.3)
    else:
        s.add_actor(ActorRec("slab", "static", [ShapeRec(SHAPE_BOX, pose7(), np.array([4.0, 1.0, 0.1]), mu=0.3)], pose7([1.5, 0, -0.1])))
    s.add_actor(ActorRec("ball", "dynamic", [ShapeRec(SHAPE_SPHERE, pose7(), np.array([r, 0, 0]), mu=0.3)], pose7([0, 0, r]), angular_damping=0.0))
    cm = s.compile()
    w = OracleWorld(cm, "f64")
    b = w.get_bodies()
    b[0, 0, 7] = v0
    w.set_bodies(b)
    w.step(150)   # sliding lasts v0 * 2 / (7 mu g) = 0.1 s; 1.5 s is long after
    b = w.get_bodies()[0, 0]
    assert b[7] == pytest.approx(5.0 / 7.0 * v0, rel=2e-2), b[7:]
    assert b[11] == pytest.approx(b[7] / r, rel=2e-2)            # omega_y = v / r: rolling without slipping
    assert abs(b[2] - r) < 1e-3 and abs(b[8]) < 1e-3 and abs(b[1]) < 1e-3
    v_before = b[7]
    w.step(50)
    assert w.get_bodies()[0, 0, 7] == pytest.approx(v_before, rel=5e-3)   # steady rolling


def test_torsional_friction_of_a_spinning_box():
    """Torsional row of a friction patch: a box spinning about the vertical on the ground is braked by the torque mu N r_eff, r_eff =
    mean distance of the patch points from their centroid (here the four bottom corners: L / sqrt 2)."""
    L, w0, mu = 0.1, 5.0, 0.3
    s = SceneDesc(1, SimParams())
    ground(s, mu=mu)
    s.add_actor(ActorRec("box", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([L / 2] * 3), mu=mu)], pose7([0, 0, L / 2]), angular_damping=0.0))
    cm = s.compile()
    w = OracleWorld(cm, "f64")
    w.step(20)  # settle
    b = w.get_bodies()
    b[0, 0, 12] = w0
    w.set_bodies(b)
    alpha = mu * G * (L / np.sqrt(2)) / (L * L / 6)   # mu m g r_eff / (m L^2 / 6)
    w.step(2)
    assert w.get_bodies()[0, 0, 12] == pytest.approx(w0 - alpha * 2 * DT, rel=3e-2)
    w.step(10)
    b = w.get_bodies()[0, 0]
    assert abs(b[12]) < 1e-3 and np.abs(b[7:10]).max() < 1e-3   # stopped, and stays where it was


def test_convex_mesh_rests_on_a_convex_mesh():
    """Multi-point convex-vs-convex manifold (the support faces of both hulls): a 48-vertex prism standing on a larger one stays put for
    600 substeps and the stack loads the ground with the weight of both (a single GJK/EPA point per pair let the upper prism spin up
    and fall off within 300 substeps)."""
    from maniskill_b200.model import ActorRec, SceneDesc, SimParams, cylinder_shape, pose7
    s = SceneDesc(2, SimParams())
    ground(s)
    up = [0.7071068, 0, -0.7071068, 0]
    s.add_actor(ActorRec("base", "dynamic", [cylinder_shape(0.12, 0.05)], pose7([0, 0, 0.0501], up)))
    s.add_actor(ActorRec("top", "dynamic", [cylinder_shape(0.05, 0.04)], pose7([0.01, 0.005, 0.1405], up)))
    cm = s.compile()
    o = OracleWorld(cm, "f64")
    o.step(600)
    b = o.rigid_body_data()
    top, base = cm.actor_rows["top"], cm.actor_rows["base"]
    assert np.abs(b[:, top, 7:]).max() < 5e-3 and np.abs(b[:, base, 7:]).max() < 5e-3
    assert abs(b[0, top, 2] - 0.14) < 2e-3 and abs(b[0, top, 0] - 0.01) < 1e-3 and abs(b[0, top, 1] - 0.005) < 1e-3
    m_top = cm.arrays["fb_mass"][cm.actor_fb["top"]]
    m_base = cm.arrays["fb_mass"][cm.actor_fb["base"]]
    imp_top = o.pair_impulse(top, base)[0]
    assert imp_top[2] == pytest.approx(m_top * 9.81 * 0.01, rel=0.03)
    imp_ground = o.pair_impulse(base, -2)[0]     # net impulse on the base: ground pushes up with both weights, the top pushes down with its own
    assert imp_ground[2] == pytest.approx(m_base * 9.81 * 0.01, rel=0.05)


def _rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def _qmat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_narrowphase_against_independent_geometry():
    """The oracle's narrowphase against computations that share nothing with it (numpy / scipy): (1) a rotated box over the ground plane --
    separations and points are the box corners below the margin; (2) a sphere near a rotated box -- closest point on the box in closed
    form; (3) two separated convex hulls (GJK) -- the distance between the two point sets' hulls from a quadratic programme over convex
    combinations (scipy SLSQP)."""
    from scipy.optimize import minimize
    rng = np.random.default_rng(11)
    plane = dict(type=SHAPE_PLANE, pose=[0, 0, 0, 0.7071068, 0, -0.7071068, 0])   # normal +z
    for _ in range(20):
        h = rng.uniform(0.02, 0.08, 3)
        q = _rand_quat(rng)
        R = _qmat(q)
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h
        zmin = (corners @ R.T)[:, 2].min()
        t = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), -zmin + rng.uniform(-0.004, 0.01)])
        world = corners @ R.T + t
        c = collide(dict(type=SHAPE_BOX, pose=list(t) + list(q), size=list(h)), plane, margin=0.02)
        below = np.sort(world[world[:, 2] < 0.02][:, 2])
        assert len(c) == min(len(below), 4) and len(c) >= 1
        assert np.allclose(np.sort(c[:, 6])[0], below[0], atol=1e-9) and np.allclose(c[:, 3:6], [0, 0, 1], atol=1e-9)
        for p, sep in zip(c[:, :3], c[:, 6]):     # every reported point sits midway between a corner and the plane, separation = the corner's height
            d = np.linalg.norm(world - (p + np.array([0, 0, 0.5 * sep])), axis=1)
            assert d.min() < 1e-9 and abs(world[d.argmin(), 2] - sep) < 1e-9
    for _ in range(20):
        h = rng.uniform(0.02, 0.08, 3)
        q = _rand_quat(rng)
        R = _qmat(q)
        r = rng.uniform(0.01, 0.04)
        local = rng.uniform(-1, 1, 3) * (h + r + 0.05)
        if np.all(np.abs(local) <= h):
            local[0] = h[0] + r + 0.005
        closest = np.clip(local, -h, h)
        dist = np.linalg.norm(local - closest)
        centre = R @ local + np.array([0.3, -0.2, 0.5])
        c = collide(dict(type=SHAPE_SPHERE, pose=list(centre) + [1, 0, 0, 0], size=[r, 0, 0]),
                    dict(type=SHAPE_BOX, pose=[0.3, -0.2, 0.5] + list(q), size=list(h)), margin=0.02)
        if dist - r < 0.02:
            assert len(c) == 1 and c[0, 6] == pytest.approx(dist - r, abs=1e-9)
            assert np.allclose(c[0, 3:6], R @ ((local - closest) / dist), atol=1e-7)     # normal from the box towards the sphere
        else:
            assert len(c) == 0
    for _ in range(12):
        A = rng.normal(size=(14, 3)) * 0.04
        B = rng.normal(size=(11, 3)) * 0.04
        ta = np.array([0.0, 0.0, 0.0])
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        # place B along d so that the hulls are separated by a gap inside the margin
        gap = rng.uniform(0.002, 0.015)
        tb = d * ((A @ d).max() - (B @ d).min() + gap)

        def f(x):
            la, mu = x[:len(A)], x[len(A):]
            v = la @ A - (mu @ B + tb)
            return v @ v
        cons = [dict(type="eq", fun=lambda x: x[:len(A)].sum() - 1), dict(type="eq", fun=lambda x: x[len(A):].sum() - 1)]
        x0 = np.concatenate([np.full(len(A), 1 / len(A)), np.full(len(B), 1 / len(B))])
        res = minimize(f, x0, bounds=[(0, 1)] * len(x0), constraints=cons, method="SLSQP", options=dict(maxiter=500, ftol=1e-16))
        dist = np.sqrt(res.fun)
        assert dist >= gap - 1e-7     # the gap between the support planes along d is a lower bound
        c = collide(dict(type=4, pose=list(ta) + [1, 0, 0, 0], size=[0, 0, 0], verts=A.astype(np.float32)),
                    dict(type=4, pose=list(tb) + [1, 0, 0, 0], size=[0, 0, 0], verts=B.astype(np.float32)), margin=0.04)
        if dist < 0.035:
            assert len(c) >= 1 and c[:, 6].min() == pytest.approx(dist, abs=2e-5), (c[:, 6], dist)


def test_double_pendulum_follows_the_lagrangian_equations():
    """Two coupled links (the off-diagonal terms of the articulated-body recursion, Coriolis / centrifugal bias forces): a planar double
    pendulum of two point masses against its textbook equations of motion integrated independently (RK4, 1e-4 s) -- large amplitude, 0.6 s."""
    l1, l2, m1, m2 = 0.4, 0.3, 1.2, 0.7
    tiny = 1e-9
    robot = dict(name="dp", links=[root_link(),
                                   link("a", 0, "revolute", (0, 0, 0), (0, 1, 0), m1, com=(0, 0, -l1), inertia=(tiny, tiny, tiny, 0, 0, 0)),
                                   link("b", 1, "revolute", (0, 0, -l1), (0, 1, 0), m2, com=(0, 0, -l2), inertia=(tiny, tiny, tiny, 0, 0, 0))],
                 disabled_collision_pairs=[])
    s = SceneDesc(1, SimParams(sim_freq=2000))
    s.add_articulation(ArticulationRec("dp", robot, pose7(), disable_gravity=False))
    cm = s.compile()
    w = OracleWorld(cm, "f64")
    q0 = np.array([1.1, -0.7])          # joint angles: the second one is relative to the first link
    w.set_joint("qpos", [q0])
    steps = 1200
    w.step(steps)
    q_sim, qd_sim = w.get_joint("qpos")[0], w.get_joint("qvel")[0]

    # rotation about +y by q takes the hanging direction -z to (-sin q, 0, -cos q): with t1 = q1, t2 = q1 + q2 measured from the downward
    # vertical (sign of x irrelevant for the planar equations)
    def f(y):
        t1, t2, w1, w2 = y
        d = t1 - t2
        den = 2 * m1 + m2 - m2 * np.cos(2 * d)
        a1 = (-G * (2 * m1 + m2) * np.sin(t1) - m2 * G * np.sin(t1 - 2 * t2) - 2 * np.sin(d) * m2 * (w2 * w2 * l2 + w1 * w1 * l1 * np.cos(d))) / (l1 * den)
        a2 = (2 * np.sin(d) * (w1 * w1 * l1 * (m1 + m2) + G * (m1 + m2) * np.cos(t1) + w2 * w2 * l2 * m2 * np.cos(d))) / (l2 * den)
        return np.array([w1, w2, a1, a2])
    y = np.array([q0[0], q0[0] + q0[1], 0.0, 0.0])
    h = 1e-4
    for _ in range(int(round(steps / 2000 / h))):
        k1 = f(y); k2 = f(y + 0.5 * h * k1); k3 = f(y + 0.5 * h * k2); k4 = f(y + h * k3)
        y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    ref_q = np.array([y[0], y[1] - y[0]])
    ref_qd = np.array([y[2], y[3] - y[2]])
    # first-order integrator at 0.5 ms against RK4: a few milliradians after 0.6 s of a swing of more than one radian
    assert np.abs(q_sim - ref_q).max() < 6e-3, (q_sim, ref_q)
    assert np.abs(qd_sim - ref_qd).max() < 5e-2, (qd_sim, ref_qd)
    assert np.abs(q_sim - q0).max() > 0.8   # it did swing


def _sat_box_box(ha, Ra, ta, hb, Rb, tb):
    """Separating-axis test of two boxes, written from the textbook statement (15 axes): -> (smallest overlap, its axis pointing from b to a, second smallest)."""
    axes = [Ra[:, i] for i in range(3)] + [Rb[:, i] for i in range(3)]
    for i in range(3):
        for j in range(3):
            c = np.cross(Ra[:, i], Rb[:, j])
            if np.linalg.norm(c) > 1e-6:
                axes.append(c / np.linalg.norm(c))
    d = ta - tb
    overlaps = []
    for n in axes:
        ra = sum(ha[i] * abs(n @ Ra[:, i]) for i in range(3))
        rb = sum(hb[i] * abs(n @ Rb[:, i]) for i in range(3))
        overlaps.append((ra + rb - abs(n @ d), n if n @ d >= 0 else -n))
    overlaps.sort(key=lambda t: t[0])
    return overlaps[0][0], overlaps[0][1], overlaps[1][0]


def test_penetration_depth_against_separating_axes():
    """Penetrating pairs, against statements that share nothing with the oracle's SAT / clipping / EPA code: (1) box-box -- the deepest reported separation is
    minus the smallest overlap over the 15 separating axes and the normal is that axis (from b to a); (2) hull-hull -- the penetration depth is the minimum over
    unit directions of h_A(d) + h_B(-d) (support functions), attained at a face normal of one hull or at a cross product of two edges: enumerated with
    scipy's ConvexHull."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(60):
        ha, hb = rng.uniform(0.02, 0.08, 3), rng.uniform(0.02, 0.08, 3)
        qa, qb = _rand_quat(rng), _rand_quat(rng)
        Ra, Rb = _qmat(qa), _qmat(qb)
        ta = np.zeros(3)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        # b slides in along d until the smallest overlap over all axes is a few millimetres (bisection on the distance, with the SAT above)
        reach = sum(ha[i] * abs(d @ Ra[:, i]) for i in range(3)) + sum(hb[i] * abs(d @ Rb[:, i]) for i in range(3))
        target, lo, hi = rng.uniform(0.001, 0.006), 0.0, reach
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if _sat_box_box(ha, Ra, ta, hb, Rb, d * mid)[0] > target:
                lo = mid
            else:
                hi = mid
        tb = d * hi
        depth, axis, second = _sat_box_box(ha, Ra, ta, hb, Rb, tb)
        if depth <= 1e-5 or second - depth < 5e-4:
            continue      # separated after all (d is not the best axis), or two axes tie: the manifold's axis is then a matter of tie-breaking
        c = collide(dict(type=SHAPE_BOX, pose=list(ta) + list(qa), size=list(ha)), dict(type=SHAPE_BOX, pose=list(tb) + list(qb), size=list(hb)), margin=0.02)
        assert 1 <= len(c) <= 4
        assert c[:, 6].min() == pytest.approx(-depth, abs=2e-6), (c[:, 6], depth)
        assert np.allclose(c[:, 3:6], axis, atol=1e-5), (c[:, 3:6], axis)
        checked += 1
    assert checked >= 25

    def depth_hulls(A, B):
        best = np.inf
        ha_, hb_ = ConvexHull(A), ConvexHull(B)
        dirs = [eq[:3] for eq in ha_.equations] + [-eq[:3] for eq in hb_.equations]     # B leaves along a face normal of A, or against one of its own
        ea = {tuple(sorted((s[i], s[(i + 1) % 3]))) for s in ha_.simplices for i in range(3)}
        eb = {tuple(sorted((s[i], s[(i + 1) % 3]))) for s in hb_.simplices for i in range(3)}
        for i, j in ea:
            for k, l in eb:
                n = np.cross(A[i] - A[j], B[k] - B[l])
                if np.linalg.norm(n) > 1e-12:
                    dirs += [n / np.linalg.norm(n), -n / np.linalg.norm(n)]
        for n in dirs:
            n = n / np.linalg.norm(n)
            best = min(best, (A @ n).max() - (B @ n).min())     # how far B must move along +n to clear A
        return best

    checked = 0
    for _ in range(10):
        A = rng.normal(size=(12, 3)) * 0.04
        B = rng.normal(size=(10, 3)) * 0.04
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        target, lo, hi = rng.uniform(0.002, 0.008), 0.0, (A @ d).max() - (B @ d).min()
        for _ in range(40):      # slide B in along d until the penetration depth is a few millimetres
            mid = 0.5 * (lo + hi)
            if depth_hulls(A, B + d * mid) > target:
                lo = mid
            else:
                hi = mid
        Bw = B + d * hi
        depth = depth_hulls(A, Bw)
        if depth < 5e-4:
            continue
        c = collide(dict(type=4, pose=[0, 0, 0, 1, 0, 0, 0], size=[0, 0, 0], verts=A.astype(np.float32)),
                    dict(type=4, pose=[0, 0, 0, 1, 0, 0, 0], size=[0, 0, 0], verts=Bw.astype(np.float32)), margin=0.04)
        assert len(c) >= 1 and c[:, 6].min() == pytest.approx(-depth, abs=5e-5), (c[:, 6], depth)
        checked += 1
    assert checked >= 6


def test_spatial_chain_follows_its_numerical_lagrangian():
    """Three-dimensional coupling of the articulated-body recursion (non-parallel axes, full inertia tensors, offset centres of mass, a prismatic joint in the
    chain: gyroscopic and Coriolis terms the planar pendulums do not have) against the Euler-Lagrange equations of the same chain built from NOTHING but its
    forward kinematics: M(q) = sum m Jv'Jv + Jw' R I R' Jw with the Jacobians taken by central differences of an FK written here, bias forces from numerical
    derivatives of M and of the potential, integrated with RK4."""
    axes = [(0, 0, 1), (0, 1, 0), (1, 0, 0), (1, 0, 0)]
    kinds = ["revolute", "revolute", "prismatic", "revolute"]
    origins = [(0, 0, 0.1), (0.05, 0, 0.2), (0.1, 0.02, -0.15), (0.2, 0, 0)]
    masses = [1.5, 1.0, 0.6, 0.8]
    coms = [(0.02, 0.01, 0.1), (0.1, -0.02, -0.05), (0.05, 0.0, 0.02), (0.03, 0.08, -0.04)]
    inertias = [(4e-3, 3e-3, 2e-3, 5e-4, -3e-4, 2e-4), (2e-3, 5e-3, 4e-3, -4e-4, 1e-4, 3e-4), (1e-3, 1.5e-3, 1.2e-3, 1e-4, 1e-4, -1e-4), (3e-3, 2e-3, 2.5e-3, 2e-4, -2e-4, 1e-4)]
    links = [root_link()] + [link(f"l{i}", i, kinds[i], origins[i], axes[i], masses[i], com=coms[i], inertia=inertias[i]) for i in range(4)]
    s = SceneDesc(1, SimParams(sim_freq=4000))
    s.add_articulation(ArticulationRec("chain", dict(name="chain", links=links, disabled_collision_pairs=[]), pose7(), disable_gravity=False))
    w = OracleWorld(s.compile(), "f64")
    q0 = np.array([0.4, 0.9, 0.05, -0.6])
    qd0 = np.array([1.5, -1.0, 0.2, 2.0])
    w.set_joint("qpos", [q0])
    w.set_joint("qvel", [qd0])
    steps = 1200                      # 0.3 s
    w.step(steps)
    q_sim, qd_sim = w.get_joint("qpos")[0], w.get_joint("qvel")[0]

    def rot(axis, a):
        x, y, z = axis
        K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=float)
        return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)

    def fk(q):
        R, p, out = np.eye(3), np.zeros(3), []
        for i in range(4):
            p = p + R @ np.array(origins[i])
            if kinds[i] == "revolute":
                R = R @ rot(axes[i], q[i])
            else:
                p = p + R @ (np.array(axes[i], dtype=float) * q[i])
            out.append((R, p + R @ np.array(coms[i])))
        return out

    def tensor(t):
        return np.array([[t[0], t[3], t[4]], [t[3], t[1], t[5]], [t[4], t[5], t[2]]])

    def mass_matrix_and_potential(q):
        h = 1e-6
        base = fk(q)
        M = np.zeros((4, 4))
        Jv = np.zeros((4, 3, 4))
        Jw = np.zeros((4, 3, 4))
        for j in range(4):
            e = np.zeros(4)
            e[j] = h
            fp, fm = fk(q + e), fk(q - e)
            for i in range(4):
                Jv[i, :, j] = (fp[i][1] - fm[i][1]) / (2 * h)
                W = ((fp[i][0] - fm[i][0]) / (2 * h)) @ base[i][0].T          # skew(omega per unit joint rate)
                Jw[i, :, j] = [W[2, 1], W[0, 2], W[1, 0]]
        V = 0.0
        for i in range(4):
            R = base[i][0]
            M += masses[i] * Jv[i].T @ Jv[i] + Jw[i].T @ (R @ tensor(inertias[i]) @ R.T) @ Jw[i]
            V += masses[i] * G * base[i][1][2]
        return M, V

    def accel(q, qd):
        h = 1e-5
        M, _ = mass_matrix_and_potential(q)
        dM, dV = [], np.zeros(4)
        for k in range(4):
            e = np.zeros(4)
            e[k] = h
            Mp, Vp = mass_matrix_and_potential(q + e)
            Mm, Vm = mass_matrix_and_potential(q - e)
            dM.append((Mp - Mm) / (2 * h))
            dV[k] = (Vp - Vm) / (2 * h)
        Mdot = sum(dM[k] * qd[k] for k in range(4))
        c = Mdot @ qd - 0.5 * np.array([qd @ dM[k] @ qd for k in range(4)])
        return np.linalg.solve(M, -c - dV)

    y = np.concatenate([q0, qd0])
    f = lambda y: np.concatenate([y[4:], accel(y[:4], y[4:])])
    h = 1e-3
    for _ in range(int(round(steps / 4000 / h))):
        k1 = f(y); k2 = f(y + 0.5 * h * k1); k3 = f(y + 0.5 * h * k2); k4 = f(y + h * k3)
        y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert np.abs(q_sim - q0).max() > 0.3                                     # it moved
    # first-order integrator at 0.25 ms against RK4 over 0.3 s
    assert np.abs(q_sim - y[:4]).max() < 1e-3, (q_sim, y[:4])          # measured 2.8e-4 rad / m
    assert np.abs(qd_sim - y[4:]).max() < 4e-3, (qd_sim, y[4:])        # measured 1.1e-3
