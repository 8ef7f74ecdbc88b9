"""gymnasium.vector.utils.batch_space: n copies of a space as one batched space."""
import numpy as np


def batch_space(space, n: int = 1):
    from ..spaces import Box, Dict, Discrete, MultiBinary, MultiDiscrete, Tuple
    if isinstance(space, Box):
        reps = (n,) + (1,) * space.low.ndim
        return Box(np.tile(space.low, reps), np.tile(space.high, reps), dtype=space.dtype)
    if isinstance(space, Discrete):
        return MultiDiscrete(np.full((n,), space.n, dtype=np.int64))
    if isinstance(space, MultiDiscrete):
        return Box(np.zeros((n,) + space.nvec.shape, dtype=space.dtype), np.tile(space.nvec - 1, (n,) + (1,) * space.nvec.ndim), dtype=space.dtype)
    if isinstance(space, MultiBinary):
        return Box(0, 1, (n,) + space.shape, dtype=space.dtype)
    if isinstance(space, Dict):
        return Dict({k: batch_space(s, n) for k, s in space.spaces.items()})
    if isinstance(space, Tuple):
        return Tuple(tuple(batch_space(s, n) for s in space.spaces))
    raise ValueError(f"cannot batch space of type {type(space)}")
