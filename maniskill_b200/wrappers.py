"""Observation wrappers the reference's learning baselines put between the env and the policy
(mani_skill/utils/wrappers/flatten.py:13-95): they turn the visual-mode observation dict into {"state", "rgb" / "depth" / "rgbd"} tensors
or one flat vector.  Pure views / concatenations on the device; the vector wrapper goes outside of them, as in the reference
(`ManiSkillVectorEnv(FlattenRGBDObservationWrapper(env))`)."""
from __future__ import annotations

import torch

from . import utils as U


class _ObservationWrapper:
    """gymnasium.ObservationWrapper: `observation()` applied to what `reset` and `step` return."""

    def __init__(self, env):
        self.env = env
        self.base_env = getattr(env, "base_env", env)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), rew, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


class FlattenRGBDObservationWrapper(_ObservationWrapper):
    """flatten.py:13-77.  rgb / depth of all cameras are concatenated along the channel axis; everything that is not sensor data is
    flattened into "state".  `sep_depth=False` merges colour and depth into one "rgbd" tensor (depth promoted to the common dtype)."""

    def __init__(self, env, rgb=True, depth=True, state=True, sep_depth=True):
        super().__init__(env)
        mode = self.base_env.obs_mode_struct
        if not mode.visual or mode.pointcloud:
            raise ValueError(f"FlattenRGBDObservationWrapper needs a camera observation mode, the env uses '{self.base_env.obs_mode}'")
        # flatten.py:36-41: textures the observation mode does not deliver are dropped from the request
        self.include_rgb, self.include_depth = rgb and mode.rgb, depth and mode.depth
        self.include_state, self.sep_depth = state, sep_depth

    def observation(self, observation: dict):
        observation = dict(observation)
        sensor_data = observation.pop("sensor_data")
        del observation["sensor_param"]
        rgb = [cam["rgb"] for cam in sensor_data.values()] if self.include_rgb else []
        depth = [cam["depth"] for cam in sensor_data.values()] if self.include_depth else []
        rgb = torch.cat(rgb, dim=-1) if rgb else None
        depth = torch.cat(depth, dim=-1) if depth else None
        ret = dict()
        if self.include_state:
            ret["state"] = U.flatten_state_dict(observation)
        if rgb is not None and depth is None:
            ret["rgb"] = rgb
        elif rgb is not None and depth is not None:
            if self.sep_depth:
                ret["rgb"], ret["depth"] = rgb, depth
            else:
                ret["rgbd"] = torch.cat([rgb, depth], dim=-1)
        elif depth is not None:
            ret["depth"] = depth
        return ret


class FlattenObservationWrapper(_ObservationWrapper):
    """flatten.py:80-95: a dict observation (e.g. obs mode "state_dict") as one vector."""

    def observation(self, observation):
        return U.flatten_state_dict(observation)
