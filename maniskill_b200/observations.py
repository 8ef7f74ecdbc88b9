"""Observation-mode parsing and the point-cloud view of the camera targets.

Mirror of mani_skill/envs/utils/observations/__init__.py:37-105 (`parse_obs_mode_to_struct`) and observations.py:16-68
(`sensor_data_to_pointcloud`).  The rasteriser writes two targets per camera (Color u8x4, PositionSegmentation i16x4 -- x, y, z in mm
in the OpenGL camera frame + segmentation id); every texture an observation mode can ask for is a slice of those, so a mode is just
a set of flags.  `normal` and `albedo` are not produced by the `minimal` shader pack this backend mirrors.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

TEXTURES = ("rgb", "depth", "segmentation", "position")
UNSUPPORTED_TEXTURES = ("normal", "albedo")


@dataclass(frozen=True)
class ObsMode:
    state_dict: bool = False
    state: bool = False
    rgb: bool = False
    depth: bool = False
    segmentation: bool = False
    position: bool = False
    pointcloud: bool = False
    raw: bool = False       # "sensor_data": every texture the shader produces

    @property
    def use_state(self) -> bool:
        """Whether privileged state (object poses ...) belongs in the observation (observations/__init__.py:27-30)."""
        return self.state or self.state_dict

    @property
    def visual(self) -> bool:
        return self.rgb or self.depth or self.segmentation or self.position


def parse_obs_mode(obs_mode: str) -> ObsMode:
    """observations/__init__.py:37-105: named modes first, otherwise a '+'-separated list of textures and state flags."""
    if obs_mode == "rgbd":
        return ObsMode(rgb=True, depth=True)
    if obs_mode == "pointcloud":
        return ObsMode(rgb=True, segmentation=True, position=True, pointcloud=True)
    if obs_mode == "sensor_data":
        return ObsMode(rgb=True, depth=True, segmentation=True, position=True, raw=True)
    parts = obs_mode.split("+")
    flags = dict(pointcloud="pointcloud" in parts)
    if flags["pointcloud"]:
        parts = [p for p in parts if p != "pointcloud"] + ["position", "rgb", "segmentation"]
    for p in parts:
        if p in ("state", "state_dict", "none"):
            continue
        if p in UNSUPPORTED_TEXTURES:
            raise NotImplementedError(f"texture '{p}' (obs mode '{obs_mode}') is not produced by this backend's shader: it writes {TEXTURES}")
        if p not in TEXTURES:
            raise NotImplementedError(f"Invalid texture type '{p}' requested in the obs mode '{obs_mode}'. Each individual texture must be one of "
                                      f"{list(TEXTURES + UNSUPPORTED_TEXTURES)}")
    return ObsMode(state_dict="state_dict" in parts, state="state" in parts, **{t: t in parts for t in TEXTURES}, **flags)


def sensor_data_to_pointcloud(observation: dict) -> dict:
    """observations.py:16-68: every camera's position texture (mm, OpenGL camera frame) becomes homogeneous world points
    `xyzw` [N, P, 4] (w = 0 for background pixels, whose segmentation id is 0), with `rgb` [N, P, 3] and `segmentation` [N, P, 1] alongside;
    the cameras of a sub-scene are concatenated along P and their entries leave `sensor_data`."""
    clouds = []
    for uid in list(observation["sensor_data"].keys()):
        images = observation["sensor_data"][uid]
        position = images["position"].float()
        position = position / 1000.0                      # a new tensor: the render target itself stays in millimetres
        seg = images["segmentation"]
        cam2world = observation["sensor_param"][uid]["cam2world_gl"].to(position.device)
        n = position.shape[0]
        xyzw = torch.cat([position, (seg != 0).to(position.dtype)], dim=-1).reshape(n, -1, 4) @ cam2world.transpose(1, 2)
        pcd = dict(xyzw=xyzw)
        if "rgb" in images:
            pcd["rgb"] = images["rgb"][..., :3].reshape(n, -1, 3).clone()
        pcd["segmentation"] = seg.reshape(n, -1, 1).clone()
        clouds.append(pcd)
        del observation["sensor_data"][uid]
    if clouds:
        observation["pointcloud"] = {k: torch.cat([c[k] for c in clouds], dim=1) for k in clouds[0]}
    return observation
