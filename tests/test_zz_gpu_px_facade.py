"""-m gpu: `scene.px` on the CUDA backend (sorted last: added after the round's GPU budget was spent, so a surprise here cannot
hide the parity tests behind `pytest -x`)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_px_facade_control_step_on_gpu():
    """`scene.px` (maniskill_b200/physx_shim.py): the reference's control-step call sequence written against `px`
    (sapien_env.py:1110-1131) leaves the same buffers as the env's own step; contact queries through `px` agree with the Scene's."""
    import maniskill_b200 as ms
    n = 8
    envs = [ms.make("PickCube-v1", num_envs=n, obs_mode="state", control_mode="pd_joint_pos", fused=False) for _ in range(2)]
    for e in envs:
        e.reset(seed=3)
    ref, raw = envs
    px = raw.scene.px
    g = torch.Generator(device=ref.device).manual_seed(0)
    for _ in range(3):
        a = 2 * torch.rand((n, 8), device=ref.device, generator=g) - 1
        ref.step(a)
        raw.agent.set_action(a)
        raw.scene._dirty = 0
        px.gpu_apply_articulation_target_position()
        for _ in range(5):
            px.step()
        px.gpu_fetch_rigid_dynamic_data()
        px.gpu_fetch_articulation_link_pose()
        for f in ("qpos", "qvel", "qacc", "target_qpos", "target_qvel"):
            getattr(px, f"gpu_fetch_articulation_{f}")()
    assert px.cuda_rigid_body_data.torch().is_cuda
    assert torch.allclose(px.cuda_rigid_body_data.torch(), ref.scene.world.rigid_body_data, atol=1e-6)
    assert torch.allclose(px.cuda_articulation_qpos.torch(), ref.scene.world.qpos, atol=1e-6)
    cubes = [[b for b in px.bodies[e] if b.name == "cube"][0] for e in range(n)]
    q = px.gpu_create_contact_body_impulse_query(cubes)
    px.gpu_query_contact_body_impulses(q)
    assert torch.allclose(q.cuda_impulses.torch(), raw.scene.get_net_contact_impulses(raw.cube), atol=1e-7)
    for e in envs:
        e.close()
