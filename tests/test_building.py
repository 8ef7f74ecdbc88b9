"""CPU: the scene-building surface (maniskill_b200/building.py) -- `sapien.Pose` algebra against scipy, and scenes assembled with the
reference's builder calls (`create_actor_builder().add_*_collision/visual ... build(name)`, `actors.build_*`) compile to the same model
tables as the direct descriptions the task mirrors use."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from maniskill_b200 import building as B
from maniskill_b200.model import SHAPE_BOX, SHAPE_CONVEX, SHAPE_SPHERE, SceneDesc
from maniskill_b200.scenes import add_table_scene, panda_articulation, pick_cube_scene


def rand_pose(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return B.Pose(rng.normal(size=3), q)


def as_rot(pose):
    w, x, y, z = pose.q.astype(np.float64)
    return Rotation.from_quat([x, y, z, w])


def test_pose_algebra_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = rand_pose(rng), rand_pose(rng)
        c = a * b
        assert np.allclose(c.p, a.p + as_rot(a).apply(b.p), atol=1e-6)
        assert np.allclose(as_rot(c).as_matrix(), as_rot(a).as_matrix() @ as_rot(b).as_matrix(), atol=1e-6)
        e = a * a.inv()
        assert np.allclose(e.p, 0, atol=1e-6) and np.allclose(np.abs(e.q), [1, 0, 0, 0], atol=1e-6)
        T = a.to_transformation_matrix()
        assert T.dtype == np.float32 and np.allclose(T[:3, :3], as_rot(a).as_matrix(), atol=1e-6) and np.allclose(T[:3, 3], a.p)
        back = B.Pose(T)
        assert np.allclose(back.p, a.p, atol=1e-6) and min(np.abs(back.q - a.q).max(), np.abs(back.q + a.q).max()) < 1e-5
        assert np.allclose(a.rpy, as_rot(a).as_euler("xyz"), atol=1e-5)     # fixed-axis x-y-z = transforms3d 'sxyz'
    p = B.Pose()
    assert p.p.dtype == np.float32 and np.array_equal(p.q, [1, 0, 0, 0])
    p.set_p([1, 2, 3])
    p.set_q([0, 1, 0, 0])
    assert np.array_equal(p.p, [1, 2, 3]) and np.array_equal(p.q, [0, 1, 0, 0])
    with pytest.raises(ValueError):
        B.Pose([1, 2])


def compiled_equal(a, b):
    assert a.scalars == b.scalars
    assert a.arrays.keys() == b.arrays.keys()
    for k in a.arrays:
        assert np.array_equal(a.arrays[k], b.arrays[k]), k
    assert a.actor_rows == b.actor_rows and a.actor_seg_id == b.actor_seg_id
    assert len(a.visuals) == len(b.visuals)


def test_pick_cube_scene_through_the_builder_calls_of_the_reference_task():
    """pick_cube.py:85-104: `actors.build_cube(..., color=[1,0,0,1], name="cube", initial_pose=Pose(p=[0,0,half]))` and
    `actors.build_sphere(..., name="goal_site", body_type="kinematic", add_collision=False, initial_pose=Pose())`."""
    direct = pick_cube_scene(3)
    built = SceneDesc(3)
    built.add_articulation(panda_articulation())
    add_table_scene(built)
    cube = B.build_cube(built, half_size=0.02, color=[1, 0, 0, 1], name="cube", initial_pose=B.Pose(p=[0, 0, 0.02]))
    goal = B.build_sphere(built, radius=0.025, color=[0, 1, 0, 1], name="goal_site", body_type="kinematic", add_collision=False, initial_pose=B.Pose())
    goal.hidden = True     # self._hidden_objects.append(self.goal_site)
    assert cube.body_type == "dynamic" and len(cube.shapes) == 1 and cube.shapes[0].collide and cube.shapes[0].visual   # twin records merged
    assert goal.shapes[0].type == SHAPE_SPHERE and not goal.shapes[0].collide
    compiled_equal(direct.compile(), built.compile())


def test_builder_records_and_errors():
    scene = SceneDesc(2)
    b = B.scene_desc_builder(scene)
    mat = B.PhysxMaterial(static_friction=2.0, dynamic_friction=2.0, restitution=0.0)
    b.add_box_collision(B.Pose(p=[0, 0, 0.1]), half_size=[0.1, 0.2, 0.3], material=mat, density=500, patch_radius=0.1, min_patch_radius=0.1)
    b.add_capsule_collision(radius=0.05, half_length=0.2)
    b.add_cylinder_collision(radius=0.05, half_length=0.1)
    b.add_sphere_visual(radius=0.3, material=B.RenderMaterial(base_color=[0, 0, 1, 1]))
    r = b.collision_records[0]
    assert (r.type, r.density, r.patch_radius, r.material.dynamic_friction) == ("box", 500, 0.1, 2.0) and np.allclose(r.scale, [0.1, 0.2, 0.3])
    assert b.collision_records[1].length == 0.2 and b.visual_records[0].radius == 0.3
    b.set_collision_groups([1, 1, 4, 0]).set_initial_pose(B.Pose(p=[0, 0, 1]))
    rec = b.build(name="thing")
    assert [s.type for s in rec.shapes] == [SHAPE_BOX, 3, SHAPE_CONVEX, SHAPE_SPHERE]
    assert rec.shapes[0].mu == 2.0 and rec.shapes[0].density == 500 and rec.shapes[0].patch_radius == 0.1 and tuple(rec.shapes[0].groups) == (1, 1, 4, 0)
    assert not rec.shapes[0].visual and not rec.shapes[3].collide and tuple(rec.shapes[3].color) == (0, 0, 1, 1)
    assert rec.shapes[2].vertices.shape == (48, 3)
    assert np.allclose(rec.initial_pose, [0, 0, 1, 1, 0, 0, 0])
    cm = scene.compile()
    m_expected = 500 * 8 * 0.1 * 0.2 * 0.3 + 1000 * (np.pi * 0.05 ** 2 * 0.4 + 4 / 3 * np.pi * 0.05 ** 3)
    cyl = 1000 * 0.2 * 0.5 * 24 * 0.05 ** 2 * np.sin(2 * np.pi / 24)      # the cooked 24-gon prism, a little less than pi r^2 h
    assert cm.arrays["fb_mass"][0] == pytest.approx(m_expected + cyl, rel=1e-6)
    with pytest.raises(RuntimeError):
        B.scene_desc_builder(scene).add_box_collision().build(name="thing")          # duplicate name
    with pytest.raises(ValueError):
        B.scene_desc_builder(scene).add_box_collision().build()                       # no name
    with pytest.raises(Exception, match="invalid physx body type"):
        B.scene_desc_builder(scene).set_physx_body_type("link")
    with pytest.raises(NotImplementedError):
        B.scene_desc_builder(scene).add_convex_collision_from_file("x.obj")
    with pytest.raises(NotImplementedError):
        B.scene_desc_builder(scene).set_scene_idxs([0])


def test_explicit_inertial_and_kinematic_build():
    scene = SceneDesc(1)
    b = B.scene_desc_builder(scene).add_box_collision(half_size=[0.1] * 3)
    b.set_mass_and_inertia(2.0, B.Pose(p=[0.01, 0, 0]), [0.1, 0.2, 0.3])
    b.build(name="a")
    B.scene_desc_builder(scene).add_box_collision(half_size=[0.1] * 3).set_mass_and_inertia(5.0, B.Pose(), [1, 1, 1]).build_kinematic(name="k")
    cm = scene.compile()
    assert cm.arrays["fb_mass"][0] == pytest.approx(2.0) and np.allclose(cm.arrays["fb_com"].reshape(-1, 3)[0], [0.01, 0, 0])
    assert np.allclose(cm.arrays["fb_inertia"].reshape(-1, 6)[0][:3], [0.1, 0.2, 0.3])
    assert cm.arrays["fb_mass"][1] == pytest.approx(1000 * 0.008)     # kinematic: the explicit inertial is ignored (actor_builder.py:156-160)


def test_twocolor_peg_helper_matches_the_task_mirror():
    from maniskill_b200.envs.lift_peg_upright import twocolor_peg_shapes
    scene = SceneDesc(1)
    c1, c2 = np.array([176, 14, 14, 255]) / 255, np.array([12, 42, 160, 255]) / 255
    rec = B.build_twocolor_peg(scene, length=0.12, width=0.025, color_1=c1, color_2=c2, name="peg", initial_pose=B.Pose(p=[0, 0, 0.1]))
    ref = twocolor_peg_shapes(0.12, 0.025, c1, c2)
    assert len(rec.shapes) == 3
    for s, r in zip(rec.shapes, ref):
        assert (s.type, s.collide, s.visual) == (r.type, r.collide, r.visual) and np.allclose(s.pose, r.pose) and np.allclose(s.size, r.size)
        if s.visual:
            assert np.allclose(s.color, r.color)
