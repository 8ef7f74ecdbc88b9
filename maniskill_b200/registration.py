"""Environment registry -- the subset of mani_skill/utils/registration.py (`register_env`, `make`) the stock tasks use."""
from __future__ import annotations

REGISTERED_ENVS = {}


def register_env(uid: str, max_episode_steps=None):
    def deco(cls):
        REGISTERED_ENVS[uid] = (cls, max_episode_steps)
        return cls
    return deco


def make(env_id: str, **kwargs):
    """gym.make(env_id, num_envs=..., obs_mode=..., control_mode=..., sim_config=...) equivalent."""
    if env_id not in REGISTERED_ENVS:
        raise KeyError(f"Env {env_id} not found in registry: {sorted(REGISTERED_ENVS)}")
    cls, max_steps = REGISTERED_ENVS[env_id]
    env = cls(**kwargs)
    if max_steps is not None:
        env.max_episode_steps = max_steps
    return env
