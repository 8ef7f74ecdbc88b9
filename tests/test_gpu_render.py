"""-m gpu: the CUDA rasteriser (through b2s_camera_group_create / b2s_render) against the CPU raster oracle on the same
body poses.  north_star: segmentation masks bit-exact.  Shapes/dtypes follow the reference's tests/test_envs.py:32-95
(rgb (128,128,3) uint8, depth (128,128,1) int16, segmentation (128,128,1) int16, camera parameter shapes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_rgbd_seg_obs_shapes_and_parity():
    import maniskill_b200 as ms
    from oracle import raster
    n = 6
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="rgb+depth+segmentation")
    env.reset(seed=3)
    g = torch.Generator(device=env.device).manual_seed(0)
    for _ in range(5):
        obs, *_ = env.step(2 * torch.rand((n, 8), device=env.device, generator=g) - 1)
    sd = obs["sensor_data"]["base_camera"]
    assert sd["rgb"].shape == (n, 128, 128, 3) and sd["rgb"].dtype == torch.uint8
    assert sd["depth"].shape == (n, 128, 128, 1) and sd["depth"].dtype == torch.int16
    assert sd["segmentation"].shape == (n, 128, 128, 1) and sd["segmentation"].dtype == torch.int16
    sp = obs["sensor_param"]["base_camera"]
    assert sp["extrinsic_cv"].shape == (n, 3, 4) and sp["intrinsic_cv"].shape == (n, 3, 3) and sp["cam2world_gl"].shape == (n, 4, 4)
    assert "agent" in obs and "extra" in obs and "obj_pose" not in obs["extra"]
    torch.cuda.synchronize()
    body = env.scene.world.body_view().cpu().numpy()
    (color, posseg), = raster.render(env._sensors.visuals, env._sensors.cams, body)
    seg_gpu = sd["segmentation"][..., 0].cpu().numpy()
    depth_gpu = sd["depth"][..., 0].cpu().numpy()
    rgb_gpu = sd["rgb"].cpu().numpy()
    assert np.array_equal(seg_gpu, posseg[..., 3]), f"{(seg_gpu != posseg[..., 3]).sum()} segmentation pixels differ"
    assert np.array_equal(depth_gpu, -posseg[..., 2])
    assert np.abs(rgb_gpu.astype(int) - color[..., :3].astype(int)).max() <= 1
    # the image is not empty: table, robot and cube are all visible
    ids = set(np.unique(seg_gpu).tolist())
    cm = env.cm
    assert cm.actor_seg_id["table-workspace"] in ids and cm.actor_seg_id["cube"] in ids and cm.link_seg_id["panda"]["panda_link0"] in ids
    assert cm.actor_seg_id["goal_site"] not in ids  # hidden object
    env.close()


def test_depth_of_table_top_is_analytic():
    """KAT: the pixel looking at the table top (z = 0 plane) must report the analytic ray/plane distance."""
    import maniskill_b200 as ms
    env = ms.make("PickCube-v1", num_envs=1, obs_mode="depth")
    env.reset(seed=0)
    obs = env.get_obs()
    d = obs["sensor_data"]["base_camera"]["depth"][0, :, :, 0].cpu().numpy().astype(float)
    eye, target = np.array([0.3, 0, 0.6]), np.array([-0.1, 0, 0.1])
    fwd = (target - eye) / np.linalg.norm(target - eye)
    # centre pixel ray = forward axis (up to half a pixel); hits z=0 at eye + t fwd, depth = t along the optical axis
    t = -eye[2] / fwd[2]
    centre = d[63:65, 63:65].mean()
    assert abs(centre - t * 1000) < 15, (centre, t * 1000)
    env.close()
