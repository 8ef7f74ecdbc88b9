"""`from gymnasium.vector.async_vector_env import AsyncVectorEnv` (mani_skill/examples/benchmarking/gpu_sim.py:17 imports it at module level; it is only
instantiated for the `--cpu-sim` arm, which this backend does not have)."""
from . import AsyncVectorEnv  # noqa: F401
