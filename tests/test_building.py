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




def _fk_world(cm_factory, qpos):
    """Link poses from the emulated device code after writing qpos (gpu_apply_articulation_qpos + gpu_update_articulation_kinematics)."""
    import torch
    from emu_world import EmuBackendWorld
    from maniskill_b200.backend import BUF_QPOS
    w = EmuBackendWorld(cm_factory)
    w.qpos[:, :len(qpos)] = torch.tensor(qpos, dtype=torch.float32)
    w.apply(BUF_QPOS)
    w.update_kinematics()
    return w.body_view()[0, :, :7].numpy().astype(np.float64)


def test_articulation_builder_joint_frames_against_analytic_fk():
    """`set_joint_properties(type, limits, pose_in_parent, pose_in_child)`: the joint frame is given on both sides and the joint moves
    about / along its x axis (sapien convention).  Link poses from the device kinematics equal root * P1 * Rx(q1) * C1^-1 * P2 * Tx(q2) * C2^-1
    for arbitrary frames, including a revolute joint anchored away from the link origin."""
    rng = np.random.default_rng(3)
    P1, C1, P2, C2 = (rand_pose(rng) for _ in range(4))
    root = B.Pose([0.3, -0.2, 0.5], [0.9238795, 0, 0, 0.3826834])
    scene = SceneDesc(1)
    ab = B.scene_desc_articulation_builder(scene)
    base = ab.create_link_builder().set_name("base")
    base.add_box_collision(half_size=[0.05] * 3)
    l1 = ab.create_link_builder(base).set_name("arm").set_joint_name("shoulder")
    l1.set_joint_properties("revolute", [[-3, 3]], pose_in_parent=P1, pose_in_child=C1, friction=0.1, damping=2.0)
    l1.add_box_collision(B.Pose(p=[0.1, 0, 0]), half_size=[0.1, 0.02, 0.02])
    l2 = ab.create_link_builder(l1).set_name("slide").set_joint_name("rail")
    l2.set_joint_properties("prismatic", [-0.5, 0.5], pose_in_parent=P2, pose_in_child=C2)
    l2.add_sphere_collision(radius=0.03)
    rec = ab.set_initial_pose(root).build(name="arm2", fix_root_link=True)
    assert [l["joint"]["type"] for l in rec.robot["links"]] == ["fixed", "revolute", "prismatic"] and "frame_offset" in rec.robot["links"][1]
    cm = scene.compile()
    assert cm.dof_names["arm2"] == ["shoulder", "rail"] and np.allclose(cm.arrays["dof_limit"].reshape(-1, 2), [[-3, 3], [-0.5, 0.5]])
    q1, q2 = 0.7, -0.15
    got = _fk_world(cm, [q1, q2])
    Rx = B.Pose([0, 0, 0], [np.cos(q1 / 2), np.sin(q1 / 2), 0, 0])
    Tx = B.Pose([q2, 0, 0])
    arm = root * P1 * Rx * C1.inv()
    slide = arm * P2 * Tx * C2.inv()
    rows = cm.link_rows["arm2"]
    for name, want in (("base", root), ("arm", arm), ("slide", slide)):
        g = got[rows[name]]
        assert np.allclose(g[:3], want.raw()[:3], atol=2e-6), name
        assert min(np.abs(g[3:] - want.raw()[3:]).max(), np.abs(g[3:] + want.raw()[3:]).max()) < 2e-6, name
    # mass of the arm link from its box at density 1000
    assert cm.arrays["dof_mass"][0] == pytest.approx(1000 * 8 * 0.1 * 0.02 * 0.02, rel=1e-6)


def test_cabinet_standin_through_the_articulation_builder():
    """The OpenCabinetDrawer stand-in assembled with create_link_builder / set_joint_properties compiles to the tables of the direct
    description (joint frame x axis = the drawers' -x sliding direction)."""
    from maniskill_b200.envs.open_cabinet_drawer import standin_cabinet
    from maniskill_b200.model import ArticulationRec, pose7
    robot, _ = standin_cabinet()
    direct = SceneDesc(2)
    d = ArticulationRec("cabinet", robot, pose7(), link_mu={L["name"]: 1.0 for L in robot["links"]}, disable_gravity=False)
    d.link_groups = {L["name"]: (1, 1, 1 << 29, 0) for L in robot["links"]}
    direct.add_articulation(d)
    built = SceneDesc(2)
    ab = B.scene_desc_articulation_builder(built)
    mat = B.PhysxMaterial(1.0, 1.0, 0.0)
    builders = []
    for L in robot["links"]:
        b = ab.create_link_builder(None if L["parent"] < 0 else builders[L["parent"]]).set_name(L["name"]).set_collision_groups([1, 1, 1 << 29, 0])
        for c in L["collisions"]:
            b.add_box_collision(B.Pose(c["p"], c["q"]), half_size=c["half_size"], material=mat)
        I = L["inertia"]
        b.set_mass_and_inertia(L["mass"], B.Pose(L["com"]), I[:3])
        J = L["joint"]
        if L["parent"] >= 0:
            # the drawer slides along -x of the cabinet: joint frame = link frame turned half a turn about z
            b.set_joint_name(J["name"]).set_joint_properties("prismatic", [J["lower"], J["upper"]], pose_in_parent=B.Pose(J["p"], [0, 0, 0, 1]),
                                                             pose_in_child=B.Pose([0, 0, 0], [0, 0, 0, 1]))
        builders.append(b)
    ab.build(name="cabinet")
    a, b_ = direct.compile(), built.compile()
    for k in a.arrays:
        assert np.allclose(a.arrays[k], b_.arrays[k], atol=1e-7), k
    assert a.dof_names == b_.dof_names and a.link_rows == b_.link_rows


def test_articulation_builder_errors():
    scene = SceneDesc(1)
    ab = B.scene_desc_articulation_builder(scene)
    root = ab.create_link_builder().set_name("root")
    with pytest.raises(ValueError):
        ab.create_link_builder()                       # second root
    child = ab.create_link_builder(root).set_name("child")
    with pytest.raises(RuntimeError, match="no joint properties"):
        ab.build(name="x")
    child.set_joint_properties("revolute", [-1, 1])
    with pytest.raises(NotImplementedError):
        ab.build(name="x", fix_root_link=False)
    ab.build(name="x")
    with pytest.raises(RuntimeError):
        ab.build(name="x")                             # duplicate name
    with pytest.raises(RuntimeError):
        child.build()


def test_collada_reader(tmp_path):
    """meshio.load_dae_parts (the Fetch robot's visual meshes are COLLADA): <triangles> and <polylist> primitives, node transforms, <unit>, Y_UP -> z-up,
    the bound material's diffuse colour."""
    from maniskill_b200.meshio import load_mesh_parts
    dae = """<?xml version="1.0"?>
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
 <asset><unit name="centimeter" meter="0.01"/><up_axis>Y_UP</up_axis></asset>
 <library_effects><effect id="fx"><profile_COMMON><technique sid="c"><phong><diffuse><color>0.2 0.4 0.6 1</color></diffuse></phong></technique></profile_COMMON></effect></library_effects>
 <library_materials><material id="mat"><instance_effect url="#fx"/></material></library_materials>
 <library_geometries><geometry id="g"><mesh>
  <source id="pos"><float_array id="pa" count="12">0 0 0  100 0 0  100 100 0  0 100 0</float_array>
   <technique_common><accessor source="#pa" count="4" stride="3"/></technique_common></source>
  <source id="nrm"><float_array id="na" count="3">0 0 1</float_array><technique_common><accessor source="#na" count="1" stride="3"/></technique_common></source>
  <vertices id="v"><input semantic="POSITION" source="#pos"/></vertices>
  <polylist material="m0" count="1"><input semantic="VERTEX" source="#v" offset="0"/><input semantic="NORMAL" source="#nrm" offset="1"/>
   <vcount>4</vcount><p>0 0 1 0 2 0 3 0</p></polylist>
  <triangles material="m0" count="1"><input semantic="VERTEX" source="#v" offset="0"/><p>0 1 2</p></triangles>
 </mesh></geometry></library_geometries>
 <library_visual_scenes><visual_scene id="s"><node id="n"><translate>0 0 50</translate>
  <instance_geometry url="#g"><bind_material><technique_common><instance_material symbol="m0" target="#mat"/></technique_common></bind_material></instance_geometry>
 </node></visual_scene></library_visual_scenes>
</COLLADA>"""
    path = tmp_path / "quad.dae"
    path.write_text(dae)
    parts = load_mesh_parts(str(path))
    assert len(parts) == 2
    (v2, f2, _), (v, f, col) = sorted(parts, key=lambda p: len(p[1]))
    assert f.shape == (2, 3) and f2.shape == (1, 3) and v.shape == (4, 3) and v2.shape == (3, 3)
    np.testing.assert_allclose(col, (0.2, 0.4, 0.6, 1.0))
    # centimetres, translated by 50 along the file's z, then (x, y, z)_yup -> (x, -z, y)_zup
    np.testing.assert_allclose(sorted(map(tuple, v.tolist())), sorted([(0, -0.5, 0), (1, -0.5, 0), (1, -0.5, 1), (0, -0.5, 1)]), atol=1e-6)
