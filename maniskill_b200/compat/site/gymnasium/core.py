"""gymnasium.core: Env and the wrapper base classes (gymnasium 0.29 behaviour)."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np


class Env:
    metadata: dict = {"render_modes": []}
    render_mode: Optional[str] = None
    spec = None
    action_space = None
    observation_space = None
    _np_random = None

    def step(self, action):
        raise NotImplementedError

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:
            self._np_random = np.random.default_rng(seed)

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def has_wrapper_attr(self, name: str) -> bool:
        return hasattr(self, name)

    def get_wrapper_attr(self, name: str):
        return getattr(self, name)

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False

    def __str__(self):
        return f"<{type(self).__name__} instance>" if self.spec is None else f"<{type(self).__name__}<{self.spec.id}>>"


class Wrapper(Env):
    def __init__(self, env: Env):
        self.env = env
        self._action_space = None
        self._observation_space = None
        self._metadata = None

    def __getattr__(self, name: str):
        if name.startswith("_") and name not in ("_np_random",):
            raise AttributeError(f"accessing private attribute '{name}' is prohibited")
        return getattr(self.env, name)

    def get_wrapper_attr(self, name: str):
        if name in self.__dict__ or hasattr(type(self), name):
            return getattr(self, name)
        return self.env.get_wrapper_attr(name)

    def has_wrapper_attr(self, name: str) -> bool:
        return name in self.__dict__ or hasattr(type(self), name) or self.env.has_wrapper_attr(name)

    @property
    def spec(self):
        return self.env.spec

    @classmethod
    def class_name(cls):
        return cls.__name__

    @property
    def action_space(self):
        return self.env.action_space if self._action_space is None else self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def observation_space(self):
        return self.env.observation_space if self._observation_space is None else self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    @property
    def metadata(self):
        return self.env.metadata if self._metadata is None else self._metadata

    @metadata.setter
    def metadata(self, value):
        self._metadata = value

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def np_random(self):
        return self.env.np_random

    @np_random.setter
    def np_random(self, value):
        self.env.np_random = value

    def step(self, action):
        return self.env.step(action)

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __str__(self):
        return f"<{type(self).__name__}{self.env}>"

    __repr__ = __str__


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return obs, self.reward(reward), terminated, truncated, info

    def reward(self, reward):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError
