"""Minimal `gymnasium` for running the reference's python layer where the real package is not installed:
the API subset mani_skill touches -- `Env`, `Wrapper` family, `spaces`, `register` / `make` / `registry`, `vector.VectorEnv`,
`vector.utils.batch_space`, `wrappers.TimeLimit` -- with gymnasium 0.29 semantics.  Activated only by maniskill_b200.compat.install() and only
when no real gymnasium can be imported."""
from . import spaces  # noqa: F401
from .core import ActionWrapper, Env, ObservationWrapper, RewardWrapper, Wrapper  # noqa: F401
from .spaces import Space  # noqa: F401
from .envs.registration import EnvSpec, make, register, registry, spec  # noqa: F401
from . import envs, vector, wrappers, utils  # noqa: F401

__version__ = "0.29.1"
