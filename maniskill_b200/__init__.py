"""maniskill_b200 -- B200-native batched rigid-body backend behind ManiSkill's BaseEnv.step()/reset() hot path.

    import maniskill_b200 as ms
    env = ms.make("PickCube-v1", num_envs=4096, obs_mode="state")
    obs, info = env.reset(seed=0)
    obs, rew, term, trunc, info = env.step(actions)

The physics runs in hand-written sm_100a CUDA (maniskill_b200/libb200sim.so, C-ABI in include/b200sim.h); there is no
CPU path.
"""
from .registration import make, register_env, REGISTERED_ENVS  # noqa: F401
from . import envs  # noqa: F401  (registers the tasks)
from .vector import ManiSkillVectorEnv  # noqa: F401
from .wrappers import FlattenObservationWrapper, FlattenRGBDObservationWrapper  # noqa: F401
from .trajectory import RecordEpisode, load_trajectories, replay_trajectory  # noqa: F401

__version__ = "0.1.0"
