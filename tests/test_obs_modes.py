"""CPU: observation modes of the BaseEnv mirror with cameras, on the emulated physics + the CPU raster oracle standing in for the
CUDA rasteriser (tests/emu_world.py).  Restates the reference's tests/test_envs.py:32-95 (obs-mode structure, shapes, dtypes) and the
mode parsing of mani_skill/envs/utils/observations/__init__.py:37-105."""
import numpy as np
import pytest
import torch

import maniskill_b200 as ms
from emu_world import EmuBackendWorld
from maniskill_b200.observations import parse_obs_mode


def test_parse_obs_mode():
    m = parse_obs_mode("rgbd")
    assert (m.rgb, m.depth, m.segmentation, m.position, m.use_state, m.visual) == (True, True, False, False, False, True)
    m = parse_obs_mode("pointcloud")
    assert m.pointcloud and m.rgb and m.segmentation and m.position and not m.depth
    m = parse_obs_mode("sensor_data")
    assert m.raw and m.rgb and m.depth and m.segmentation and m.position
    m = parse_obs_mode("state+rgb+segmentation")
    assert m.state and not m.state_dict and m.rgb and m.segmentation and not m.depth and m.use_state
    m = parse_obs_mode("state_dict+depth")
    assert m.state_dict and not m.state and m.depth
    for name in ("state", "state_dict", "none"):
        assert not parse_obs_mode(name).visual
    assert parse_obs_mode("state").state and parse_obs_mode("state_dict").state_dict and not parse_obs_mode("none").use_state
    with pytest.raises(NotImplementedError, match="Invalid texture type 'rgbx'"):
        parse_obs_mode("rgbx+depth")
    with pytest.raises(NotImplementedError, match="normal"):
        parse_obs_mode("rgb+normal")


@pytest.mark.parametrize("mode,textures", [("rgb", {"rgb"}), ("rgbd", {"rgb", "depth"}), ("depth+segmentation", {"depth", "segmentation"}),
                                            ("rgb+position", {"rgb", "position"}), ("sensor_data", {"rgb", "depth", "segmentation", "position"})])
def test_texture_modes_deliver_exactly_the_requested_textures(mode, textures):
    n = 2
    env = ms.make("PickCube-v1", num_envs=n, obs_mode=mode, world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    assert set(obs) == {"agent", "extra", "sensor_param", "sensor_data"} and "obj_pose" not in obs["extra"]
    sd = obs["sensor_data"]["base_camera"]
    assert set(sd) == textures
    spec = dict(rgb=(3, torch.uint8), depth=(1, torch.int16), segmentation=(1, torch.int16), position=(3, torch.int16))
    for k in textures:
        assert sd[k].shape == (n, 128, 128, spec[k][0]) and sd[k].dtype == spec[k][1]
    sp = obs["sensor_param"]["base_camera"]
    assert sp["extrinsic_cv"].shape == (n, 3, 4) and sp["intrinsic_cv"].shape == (n, 3, 3) and sp["cam2world_gl"].shape == (n, 4, 4)
    if {"depth", "position"} <= textures:
        assert torch.equal(sd["depth"][..., 0], -sd["position"][..., 2])       # depth = -z of the OpenGL camera frame (shaders.py:74-83)


def test_state_plus_textures_and_the_hidden_goal():
    env = ms.make("PickCube-v1", num_envs=2, obs_mode="state+rgb+segmentation", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    assert set(obs) == {"state", "sensor_param", "sensor_data"} and obs["state"].shape == (2, 42)      # agent / extra folded into `state`
    ids = set(np.unique(obs["sensor_data"]["base_camera"]["segmentation"].numpy()).tolist())
    cm = env.cm
    assert cm.actor_seg_id["cube"] in ids and cm.actor_seg_id["table-workspace"] in ids and cm.actor_seg_id["goal_site"] not in ids
    o2, *_ = env.step(torch.zeros(2, 8))
    assert set(o2) == set(obs)


def test_pointcloud_mode_points_lie_on_the_scene():
    """obs mode "pointcloud" (sapien_env.py:525-527): xyzw [N, H*W, 4] in the world frame.  Points with the table's id lie in the
    table-top plane z = 0, points with the cube's id inside the cube's box, background points have w = 0."""
    n = 2
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="pointcloud", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=1)
    assert obs["sensor_data"] == {} and set(obs["pointcloud"]) == {"xyzw", "rgb", "segmentation"}
    pc = obs["pointcloud"]
    P = 128 * 128
    assert pc["xyzw"].shape == (n, P, 4) and pc["rgb"].shape == (n, P, 3) and pc["segmentation"].shape == (n, P, 1)
    seg = pc["segmentation"][..., 0]
    cm = env.cm
    table = seg == cm.actor_seg_id["table-workspace"]
    assert table.sum() > 1000
    assert pc["xyzw"][..., 2][table].abs().max() < 2e-3                       # mm-quantised positions
    assert (pc["xyzw"][..., 3][seg != 0] == 1).all() and (pc["xyzw"][..., 3][seg == 0] == 0).all()
    cube = seg == cm.actor_seg_id["cube"]
    for e in range(n):
        pts = pc["xyzw"][e][cube[e]][:, :3]
        assert len(pts) > 10
        assert ((pts - env.cube.pose.p[e]).norm(dim=1) < 0.02 * 3 ** 0.5 + 2e-3).all()
    # stepping keeps the structure
    o2, *_ = env.step(torch.zeros(n, 8))
    assert set(o2["pointcloud"]) == {"xyzw", "rgb", "segmentation"}


def test_depth_of_the_table_top_is_analytic_on_the_raster_oracle():
    """KAT of the raster oracle itself (the checker of the CUDA rasteriser): the centre pixel looks along the optical axis and must
    report the analytic ray / plane distance to the table top (z = 0)."""
    env = ms.make("PickCube-v1", num_envs=1, obs_mode="depth+segmentation", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    d = obs["sensor_data"]["base_camera"]["depth"][0, :, :, 0].numpy().astype(float)
    seg = obs["sensor_data"]["base_camera"]["segmentation"][0, :, :, 0].numpy()
    eye, target = np.array([0.3, 0, 0.6]), np.array([-0.1, 0, 0.1])
    fwd = (target - eye) / np.linalg.norm(target - eye)
    t = -eye[2] / fwd[2]
    assert abs(d[63:65, 63:65].mean() - t * 1000) < 15
    # rows further down the image look at nearer table points: over the table's pixels depth decreases monotonically down a column
    on_table = seg[:, 64] == env.cm.actor_seg_id["table-workspace"]
    assert on_table.sum() > 30 and (np.diff(d[:, 64][on_table]) <= 0).all()


def test_wrist_camera_follows_its_mount():
    """PegInsertionSide-v1 (panda_wristcam): the hand camera's extrinsics move with the hand link, the base camera's do not."""
    env = ms.make("PegInsertionSide-v1", num_envs=2, obs_mode="rgbd", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=0)
    assert set(obs["sensor_data"]) == {"base_camera", "hand_camera"}
    e0 = {k: obs["sensor_param"][k]["extrinsic_cv"].clone() for k in obs["sensor_param"]}
    a = torch.zeros(2, 8)
    a[:, 1] = 1.0
    for _ in range(3):
        obs, *_ = env.step(a)
    assert torch.equal(obs["sensor_param"]["base_camera"]["extrinsic_cv"], e0["base_camera"])
    assert (obs["sensor_param"]["hand_camera"]["extrinsic_cv"] - e0["hand_camera"]).abs().max() > 1e-3
    # the hand camera sees the gripper fingers
    ids = set(np.unique(obs["sensor_data"]["hand_camera"]["segmentation"].numpy()).tolist()) if "segmentation" in obs["sensor_data"]["hand_camera"] else None
    assert ids is None      # rgbd carries no segmentation


def test_flatten_wrappers_on_an_env():
    """flatten.py:13-95 in place: PPO-RGB style observation from PegInsertionSide-v1 (two cameras), under the vector wrapper."""
    from maniskill_b200.wrappers import FlattenObservationWrapper, FlattenRGBDObservationWrapper
    env = ms.make("PegInsertionSide-v1", num_envs=2, obs_mode="rgbd", world_factory=EmuBackendWorld)
    venv = ms.ManiSkillVectorEnv(FlattenRGBDObservationWrapper(env, rgb=True, depth=True, state=True), auto_reset=True)
    obs, _ = venv.reset(seed=0)
    assert set(obs) == {"state", "rgb", "depth"}
    assert obs["rgb"].shape == (2, 128, 128, 6) and obs["rgb"].dtype == torch.uint8 and obs["depth"].shape == (2, 128, 128, 2)
    assert obs["state"].shape == (2, 9 + 9 + 7)       # qpos, qvel, tcp pose: no privileged object poses in a visual mode
    o2, r, te, tr, info = venv.step(torch.zeros(2, 8))
    assert set(o2) == set(obs) and r.shape == (2,)
    merged = FlattenRGBDObservationWrapper(ms.make("PickCube-v1", num_envs=1, obs_mode="rgb", world_factory=EmuBackendWorld), sep_depth=False)
    o, _ = merged.reset(seed=0)
    assert set(o) == {"state", "rgb"} and o["rgb"].shape == (1, 128, 128, 3)       # depth is not in the mode: dropped from the request
    with pytest.raises(ValueError):
        FlattenRGBDObservationWrapper(ms.make("PickCube-v1", num_envs=1, obs_mode="state", world_factory=EmuBackendWorld))
    flat = FlattenObservationWrapper(ms.make("PickCube-v1", num_envs=2, obs_mode="state_dict", world_factory=EmuBackendWorld))
    o, _ = flat.reset(seed=0)
    assert o.shape == (2, 42)


def test_render_modes():
    """sapien_env.py:1369-1439: `render()` under render_mode "rgb_array" (human render camera, 512 x 512, hidden objects shown),
    "sensors" (what the agent's cameras see: rgb + depth visualisation tiled) and "all"; no render mode -> RuntimeError."""
    env = ms.make("PickCube-v1", num_envs=2, obs_mode="state", render_mode="rgb_array", world_factory=EmuBackendWorld,
                  sensor_configs=dict(base_camera=dict(width=64, height=64)))
    env.reset(seed=0)
    img = env.render()
    assert img.shape == (2, 512, 512, 3) and img.dtype == torch.uint8
    # the goal site (green sphere) is hidden from the sensors but shown to the human camera
    green = (img[..., 1].int() - img[..., 0].int() > 80) & (img[..., 1].int() - img[..., 2].int() > 80)
    assert green.view(2, -1).sum(1).max() > 20        # (in some layouts the arm occludes it)
    sens = env.get_sensor_images()["base_camera"]
    assert set(sens) == {"rgb", "depth"} and sens["depth"].shape == (2, 64, 64, 3) and sens["depth"].dtype == torch.uint8
    sg = (sens["rgb"][..., 1].int() - sens["rgb"][..., 0].int() > 80) & (sens["rgb"][..., 1].int() - sens["rgb"][..., 2].int() > 80)
    assert sg.sum() == 0
    env.render_mode = "sensors"
    assert env.render().shape == (2, 64, 128, 3)
    env.render_mode = "all"
    assert env.render().shape == (2, 512, 512 + 64, 3)          # the two 64 x 64 sensor images share a column next to the 512 x 512 view
    env.render_mode = None
    with pytest.raises(RuntimeError, match="render_mode is not set"):
        env.render()
    with pytest.raises(NotImplementedError):
        ms.make("PickCube-v1", num_envs=1, render_mode="human", world_factory=EmuBackendWorld)
    for task in ("PickCube-v1", "PegInsertionSide-v1", "OpenCabinetDrawer-v1"):
        e = ms.make(task, num_envs=1, obs_mode="state", render_mode="rgb_array", world_factory=EmuBackendWorld)
        e.reset(seed=0)
        im = e.render()
        assert im.shape == (1, 512, 512, 3) and len(torch.unique(im.reshape(-1, 3), dim=0)) > 20, task
