"""`from gymnasium.vector.sync_vector_env import SyncVectorEnv`: see async_vector_env.py."""
from . import SyncVectorEnv  # noqa: F401
