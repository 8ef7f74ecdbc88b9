"""gymnasium.envs.registration: EnvSpec / WrapperSpec / register / make / registry with the wrapper order of gymnasium 0.29
(`make`: OrderEnforcing when `order_enforce`, then TimeLimit from `max_episode_steps`, then the spec's `additional_wrappers`)."""
from __future__ import annotations

import copy
import importlib
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Union


@dataclass
class WrapperSpec:
    name: str
    entry_point: str
    kwargs: Optional[dict] = None


@dataclass
class EnvSpec:
    id: str
    entry_point: Union[Callable, str, None] = None
    reward_threshold: Optional[float] = None
    nondeterministic: bool = False
    max_episode_steps: Optional[int] = None
    order_enforce: bool = True
    autoreset: bool = False
    disable_env_checker: bool = False
    apply_api_compatibility: bool = False
    kwargs: dict = field(default_factory=dict)
    additional_wrappers: tuple = ()
    vector_entry_point: Union[Callable, str, None] = None

    @property
    def name(self):
        return self.id.rsplit("-v", 1)[0]

    @property
    def version(self):
        tail = self.id.rsplit("-v", 1)
        return int(tail[1]) if len(tail) == 2 and tail[1].isdigit() else None

    def make(self, **kwargs):
        return make(self, **kwargs)


registry: dict = {}


def register(id: str, entry_point=None, reward_threshold=None, nondeterministic=False, max_episode_steps=None, order_enforce=True, autoreset=False,
             disable_env_checker=False, apply_api_compatibility=False, additional_wrappers=(), vector_entry_point=None, kwargs=None, **more):
    kw = dict(kwargs or {})
    kw.update(more)
    registry[id] = EnvSpec(id=id, entry_point=entry_point, reward_threshold=reward_threshold, nondeterministic=nondeterministic,
                           max_episode_steps=max_episode_steps, order_enforce=order_enforce, autoreset=autoreset, disable_env_checker=disable_env_checker,
                           apply_api_compatibility=apply_api_compatibility, kwargs=kw, additional_wrappers=tuple(additional_wrappers),
                           vector_entry_point=vector_entry_point)


def spec(env_id: str) -> EnvSpec:
    if env_id not in registry:
        raise KeyError(f"No registered env with id: {env_id}")
    return registry[env_id]


def load_env_creator(name: str):
    mod, attr = name.split(":")
    return getattr(importlib.import_module(mod), attr)


def make(id, max_episode_steps: Optional[int] = None, autoreset: Optional[bool] = None, apply_api_compatibility=None, disable_env_checker=None, **kwargs):
    env_spec = id if isinstance(id, EnvSpec) else spec(id)
    creator = env_spec.entry_point
    if creator is None:
        raise RuntimeError(f"{env_spec.id} registered but entry_point is not specified")
    if isinstance(creator, str):
        creator = load_env_creator(creator)
    kw = copy.deepcopy(env_spec.kwargs)
    kw.update(kwargs)
    env = creator(**kw)
    spec_ = copy.deepcopy(env_spec)
    spec_.kwargs = kw
    if max_episode_steps is not None:
        spec_.max_episode_steps = max_episode_steps
    try:
        env.unwrapped.spec = spec_
    except Exception:
        pass
    from ..wrappers import OrderEnforcing, TimeLimit
    if spec_.order_enforce:
        env = OrderEnforcing(env)
    if spec_.max_episode_steps is not None:
        env = TimeLimit(env, spec_.max_episode_steps)
    for w in spec_.additional_wrappers:
        env = load_env_creator(w.entry_point)(env=env, **(w.kwargs or {}))
    return env
