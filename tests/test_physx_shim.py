"""CPU: the `px.*` facade (maniskill_b200/physx_shim.py) over the emulated backend -- the call sequence of the
reference's control step (mani_skill/envs/sapien_env.py:1110-1131, scene.py:950-986) written against `px` gives the same buffers as
the BaseEnv mirror, the buffers alias, and the contact queries return what the mirror's Scene returns."""
import pytest
import torch

import maniskill_b200 as ms
from emu_world import EmuBackendWorld
from maniskill_b200.physx_shim import PhysxGpuSystem


def make_pair(n=3, seed=0):
    envs = [ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode="pd_joint_pos", world_factory=EmuBackendWorld, fused=False) for _ in range(2)]
    for e in envs:
        e.reset(seed=seed)
    return envs


def px_for(env):
    px = env.scene.px       # scene.py:61-63 `self.px`
    assert isinstance(px, PhysxGpuSystem) and env.scene.px is px
    return px


def test_control_step_through_px_matches_the_env_mirror():
    ref, raw = make_pair()
    px = px_for(raw)
    assert px.timestep == pytest.approx(0.01)
    px.timestep = 0.01
    with pytest.raises(RuntimeError):
        px.timestep = 0.005
    g = torch.Generator().manual_seed(1)
    for _ in range(4):
        action = 2 * torch.rand(3, 8, generator=g) - 1
        ref.step(action)
        # sapien_env.py:1110-1131 by hand on the second world: controller math, target upload, 5 x px.step(), fetch everything
        raw.agent.set_action(action)
        raw.scene._dirty = 0                                     # the mirror's dirty-mask bookkeeping is bypassed: px calls apply explicitly
        px.gpu_apply_articulation_target_position()
        for _ in range(5):
            px.step()
        px.gpu_fetch_rigid_dynamic_data()
        px.gpu_fetch_articulation_link_pose()
        px.gpu_fetch_articulation_link_velocity()
        for f in ("qpos", "qvel", "qacc", "target_qpos", "target_qvel"):
            getattr(px, f"gpu_fetch_articulation_{f}")()
        assert torch.equal(px.cuda_rigid_body_data.torch(), ref.scene.world.rigid_body_data)
        assert torch.equal(px.cuda_articulation_qpos.torch(), ref.scene.world.qpos)
        assert torch.equal(px.cuda_articulation_qvel.torch(), ref.scene.world.qvel)
        assert torch.equal(px.cuda_articulation_target_qpos.torch(), ref.scene.world.target_qpos)
    assert px.cuda_rigid_body_data.torch().data_ptr() == raw.scene.world.rigid_body_data.data_ptr()   # aliases, not copies


def test_state_set_through_px_buffers_and_indices():
    _, raw = make_pair()
    px = px_for(raw)
    N, R = raw.num_envs, raw.scene.world.n_rows
    cube = [b for b in px.bodies[1] if b.name == "cube"][0]
    assert cube.gpu_pose_index == 1 * R + raw.cube.row and px.articulations[2][0].gpu_index == 2 and px.articulations[0][0].dof == 9
    data = px.cuda_rigid_body_data.torch()
    data[cube.gpu_pose_index, :3] = torch.tensor([0.05, -0.07, 0.3])
    data[cube.gpu_pose_index, 7:] = 0
    px.gpu_apply_rigid_dynamic_data()
    q = px.cuda_articulation_qpos.torch()
    q[px.articulations[1][0].gpu_index, 0] = 0.3
    px.gpu_apply_articulation_qpos()
    px.gpu_update_articulation_kinematics()
    px.gpu_fetch_rigid_dynamic_data()
    px.gpu_fetch_articulation_qpos()
    assert torch.allclose(raw.cube.pose.p[1], torch.tensor([0.05, -0.07, 0.3])) and not torch.allclose(raw.cube.pose.p[0], raw.cube.pose.p[1])
    assert float(raw.agent.robot.get_qpos()[1, 0]) == pytest.approx(0.3)
    # the robot moved in env 1 only
    tcp = raw.agent.tcp.pose.p
    assert not torch.allclose(tcp[1], tcp[0], atol=1e-3)
    px.step()
    px.gpu_fetch_rigid_dynamic_data()
    assert float(raw.cube.pose.p[1, 2]) < 0.3 and float(raw.cube.linear_velocity[1, 2]) < -0.05   # the lifted cube falls


def test_contact_queries_through_px():
    _, raw = make_pair()
    px = px_for(raw)
    for _ in range(3):
        raw.step(torch.zeros(3, 8))
    by_name = lambda e, n: [b for b in px.bodies[e] if b.name == n][0]
    # net impulse on the resting cube of every env = m g dt upwards; pair query cube <-> table the same; cube <-> finger zero
    body_q = px.gpu_create_contact_body_impulse_query([by_name(e, "cube") for e in range(3)])
    px.gpu_query_contact_body_impulses(body_q)
    net = body_q.cuda_impulses.torch()
    assert net.shape == (3, 3)
    assert torch.allclose(net, raw.scene.get_net_contact_impulses(raw.cube), atol=1e-7)
    assert (net[:, 2] > 0).all() and torch.allclose(net[:, :2], torch.zeros(3, 2), atol=1e-5)
    pairs = [(by_name(e, "cube"), by_name(e, "table-workspace")) for e in (2, 0)] + [(by_name(1, "cube"), by_name(1, "panda_panda_leftfinger"))]
    pair_q = px.gpu_create_contact_pair_impulse_query(pairs)
    px.gpu_query_contact_pair_impulses(pair_q)
    out = pair_q.cuda_impulses.torch()
    assert out.shape == (3, 3)
    assert torch.allclose(out[0], net[2], atol=1e-7) and torch.allclose(out[1], net[0], atol=1e-7) and torch.equal(out[2], torch.zeros(3))
    assert pair_q.cuda_impulses.torch().data_ptr() == out.data_ptr()
    with pytest.raises(RuntimeError):
        px.gpu_create_contact_pair_impulse_query([(by_name(0, "cube"), by_name(1, "table-workspace"))])
