"""-m gpu: the `sapien` shim (maniskill_b200/compat) on the REAL backend -- scenes built through the sapien-shaped API (the calls the
reference's builders make), compiled at `px.gpu_init()` into one batched world through the C-ABI, stepped and rendered on the B200.
(The unmodified reference itself needs its checkout, which the GPU box does not have: tests/test_reference_unmodified.py runs it on the
emulated world where the checkout exists.)"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

URDF = """<?xml version="1.0"?>
<robot name="arm2">
  <link name="base"><inertial><mass value="1"/><inertia ixx="0.01" iyy="0.01" izz="0.01" ixy="0" ixz="0" iyz="0"/></inertial>
    <collision><origin xyz="0 0 0.05"/><geometry><box size="0.1 0.1 0.1"/></geometry></collision>
    <visual><origin xyz="0 0 0.05"/><geometry><box size="0.1 0.1 0.1"/></geometry></visual></link>
  <link name="upper"><inertial><origin xyz="0 0 0.15"/><mass value="0.5"/><inertia ixx="0.004" iyy="0.004" izz="0.0005" ixy="0" ixz="0" iyz="0"/></inertial>
    <collision><origin xyz="0 0 0.15"/><geometry><box size="0.04 0.04 0.3"/></geometry></collision>
    <visual><origin xyz="0 0 0.15"/><geometry><box size="0.04 0.04 0.3"/></geometry></visual></link>
  <link name="fore"><inertial><origin xyz="0 0 0.1"/><mass value="0.3"/><inertia ixx="0.001" iyy="0.001" izz="0.0002" ixy="0" ixz="0" iyz="0"/></inertial>
    <collision><origin xyz="0 0 0.1"/><geometry><sphere radius="0.03"/></geometry></collision></link>
  <joint name="shoulder" type="revolute"><parent link="base"/><child link="upper"/><origin xyz="0 0 0.1"/><axis xyz="0 1 0"/>
    <limit lower="-1.5" upper="1.5" effort="50" velocity="3"/></joint>
  <joint name="elbow" type="revolute"><parent link="upper"/><child link="fore"/><origin xyz="0 0 0.3"/><axis xyz="0 1 0"/>
    <limit lower="-2" upper="2" effort="50" velocity="3"/></joint>
</robot>
"""


@pytest.fixture(scope="module")
def sapien():
    import maniskill_b200.compat as compat
    compat.WORLD_FACTORY = None
    compat.install()
    import sapien
    assert sapien.__file__.startswith(compat.SITE)
    return sapien


def _scenes(sapien, n):
    px = sapien.physx.PhysxGpuSystem(device="cuda:0")
    px.timestep = 0.01
    scenes = []
    for i in range(n):
        s = sapien.Scene([px, sapien.render.RenderSystem("cuda:0")])
        px.set_scene_offset(s, [i * 5.0, 0, 0])
        scenes.append(s)
    return px, scenes


def test_actors_built_through_the_sapien_api_fall_and_rest(sapien):
    n = 12
    px, scenes = _scenes(sapien, n)
    boxes = []
    for i, s in enumerate(scenes):
        g = s.create_actor_builder()
        g.add_plane_collision(sapien.Pose(p=[0, 0, 0], q=[0.7071068, 0, -0.7071068, 0]))
        g.add_plane_visual(sapien.Pose(p=[0, 0, 0], q=[0.7071068, 0, -0.7071068, 0]), material=[0.4, 0.4, 0.4])
        g.set_physx_body_type("static")
        g.build(name=f"scene-{i}_ground")
        b = s.create_actor_builder()
        h = 0.02 + 0.002 * i                       # per-sub-scene geometry: different box sizes
        b.add_box_collision(half_size=[h, h, h])
        b.add_box_visual(half_size=[h, h, h], material=[1, 0, 0])
        b.set_initial_pose(sapien.Pose(p=[0.01 * i, 0, 0.3]))
        boxes.append(b.build(name=f"scene-{i}_box"))
    px.gpu_init()
    comp = [e.find_component_by_type(sapien.physx.PhysxRigidDynamicComponent) for e in boxes]
    assert [c.gpu_pose_index for c in comp] == list(range(n))      # one row per sub-scene
    for _ in range(150):
        px.step()
    px.gpu_fetch_rigid_dynamic_data()
    data = px.cuda_rigid_body_data.torch()
    assert data.is_cuda and data.shape == (n, 13)
    z = data[:, 2].cpu().numpy()
    want = 0.02 + 0.002 * np.arange(n)
    assert np.abs(z - want).max() < 2e-3, z                          # each box rests on ITS half height
    assert np.abs(data[:, 0].cpu().numpy() - 0.01 * np.arange(n)).max() < 6e-3   # and started at its own initial pose (it may slide a few mm on landing)
    assert float(data[:, 7:].abs().max()) < 5e-2
    # contact impulse query between each box and the ground: m g dt
    q = px.gpu_create_contact_body_impulse_query(comp)
    px.gpu_query_contact_body_impulses(q)
    imp = q.cuda_impulses.torch()[:, 2].cpu().numpy()
    mass = 1000.0 * (2 * want) ** 3
    assert np.allclose(imp, mass * 9.81 * 0.01, rtol=0.05), (imp, mass * 9.81 * 0.01)
    # render through the camera-group surface
    cams = []
    for s in scenes:
        c = sapien.render.RenderCameraComponent(64, 64)
        c.set_fovy(1.2)
        c.near, c.far = 0.01, 10.0
        e = sapien.Entity()
        e.add_component(c)
        s.add_entity(e)
        from transforms3d.euler import euler2quat
        c.local_pose = sapien.Pose(p=[-0.5, 0, 0.3], q=euler2quat(0, 0.5, 0))
        cams.append(c)
    grp = sapien.render.RenderSystemGroup([s.render_system for s in scenes])
    grp.set_cuda_poses(px.cuda_rigid_body_data)
    cg = grp.create_camera_group(cams, ["Color", "PositionSegmentation"])
    cg.take_picture()
    seg = cg.get_picture_cuda("PositionSegmentation").torch()[..., 3]
    rgb = cg.get_picture_cuda("Color").torch()
    assert seg.shape == (n, 64, 64) and rgb.shape == (n, 64, 64, 4) and rgb.dtype == torch.uint8
    box_id = boxes[0].per_scene_id
    counts = (seg == box_id).sum(dim=(1, 2)).cpu().numpy()
    assert counts.min() > 0 and counts[-1] > counts[0]              # bigger boxes cover more pixels
    red = rgb[seg == box_id]
    assert int(red[:, 0].float().mean()) > 60 and int(red[:, 1].max()) == 0


def test_urdf_loader_articulation_drive_on_cuda(sapien, tmp_path):
    path = tmp_path / "arm2.urdf"
    path.write_text(URDF)
    n = 6
    px, scenes = _scenes(sapien, n)
    arts = []
    for i, s in enumerate(scenes):
        loader = s.create_urdf_loader()
        loader.fix_root_link = True
        ab = loader.load_file_as_articulation_builder(str(path))
        ab.set_initial_pose(sapien.Pose(p=[0, 0, 0]))
        for lb in ab.link_builders:
            lb.set_name(lb.name)
        art = ab.build(fix_root_link=True, name_prefix=f"scene-{i}-arm2_")
        art.name = f"scene-{i}_arm2"
        for j in art.get_active_joints():
            j.set_drive_properties(200.0, 20.0, 50.0, "force")
        for l in art.links:
            l.disable_gravity = True
        arts.append(art)
    assert arts[0].dof == 2 and [j.name for j in arts[0].active_joints] == ["scene-0-arm2_shoulder", "scene-0-arm2_elbow"]
    px.gpu_init()
    assert [a.gpu_index for a in arts] == list(range(n))
    tq = px.cuda_articulation_target_qpos.torch()
    target = torch.tensor([[0.5, -0.8]], device=tq.device).repeat(n, 1) * torch.linspace(0.5, 1.0, n, device=tq.device)[:, None]
    tq[:, :2] = target
    px.gpu_apply_articulation_target_position()
    for _ in range(300):
        px.step()
    px.gpu_fetch_articulation_qpos()
    px.gpu_fetch_articulation_link_pose()
    q = px.cuda_articulation_qpos.torch()[:, :2]
    assert float((q - target).abs().max()) < 2e-2, (q, target)
    # link poses follow: the forearm origin sits at the end of the rotated upper arm
    fore = arts[0].links[2]
    row = px.cuda_rigid_body_data.torch()[fore.gpu_pose_index]
    a = float(q[0, 0])
    assert abs(float(row[0]) - 0.3 * np.sin(a)) < 5e-3 and abs(float(row[2]) - (0.1 + 0.3 * np.cos(a))) < 5e-3
