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


def test_peg_insertion_rgbd_two_cameras_heterogeneous_envs():
    """BASELINE.json configs[2]: PegInsertionSide-v1 with 128x128 RGBD, base + wrist camera, per-env peg/box geometry."""
    import maniskill_b200 as ms
    from oracle import raster
    n = 5
    env = ms.make("PegInsertionSide-v1", num_envs=n, obs_mode="rgbd", sensor_outputs="raw")   # raw render targets: get_picture_cuda below
    obs, _ = env.reset(seed=1)
    g = torch.Generator(device=env.device).manual_seed(0)
    for _ in range(3):
        obs, rew, term, trunc, info = env.step(2 * torch.rand((n, 8), device=env.device, generator=g) - 1)
    assert set(obs["sensor_data"].keys()) == {"base_camera", "hand_camera"}
    for cam in ("base_camera", "hand_camera"):
        assert obs["sensor_data"][cam]["rgb"].shape == (n, 128, 128, 3)
        assert obs["sensor_data"][cam]["depth"].shape == (n, 128, 128, 1)
        assert "segmentation" not in obs["sensor_data"][cam]
    assert obs["extra"]["tcp_pose"].shape == (n, 7) and torch.isfinite(rew).all()
    torch.cuda.synchronize()
    body = env.scene.world.body_view().cpu().numpy()
    ref = raster.render(env._sensors.visuals, env._sensors.cams, body)
    grp = env._sensors.group
    for i, (color, posseg) in enumerate(ref):
        ps = grp.get_picture_cuda("PositionSegmentation", i).cpu().numpy()
        col = grp.get_picture_cuda("Color", i).cpu().numpy()
        assert np.array_equal(ps[..., 3], posseg[..., 3]), f"camera {i}: {(ps[..., 3] != posseg[..., 3]).sum()} seg pixels differ"
        assert np.array_equal(ps[..., :3], posseg[..., :3])
        assert np.abs(col.astype(int) - color.astype(int)).max() <= 1
    # pegs have different lengths in different envs, and the peg is visible from the base camera
    peg_seg = env.cm.actor_seg_id["peg"]
    counts = [(grp.get_picture_cuda("PositionSegmentation", 0)[e, ..., 3] == peg_seg).sum().item() for e in range(n)]
    assert min(counts) > 0
    env.close()
