"""gymnasium.spaces.utils: flatdim / flatten_space / flatten for Box, Discrete and (nested) Dict / Tuple spaces."""
import numpy as np


def flatdim(space) -> int:
    from . import Box, Dict, Discrete, MultiBinary, MultiDiscrete, Tuple
    if isinstance(space, (Box, MultiBinary)):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    if isinstance(space, MultiDiscrete):
        return int(np.sum(space.nvec))
    if isinstance(space, Dict):
        return sum(flatdim(s) for s in space.spaces.values())
    if isinstance(space, Tuple):
        return sum(flatdim(s) for s in space.spaces)
    raise NotImplementedError(type(space))


def flatten_space(space):
    from . import Box, Dict, Discrete, Tuple
    if isinstance(space, Box):
        return Box(space.low.flatten(), space.high.flatten(), dtype=space.dtype)
    if isinstance(space, Discrete):
        return Box(0, 1, (space.n,), dtype=space.dtype)
    if isinstance(space, (Dict, Tuple)):
        subs = [flatten_space(s) for s in (space.spaces.values() if isinstance(space, Dict) else space.spaces)]
        dtype = np.result_type(*[s.dtype for s in subs]) if subs else np.float32
        return Box(np.concatenate([s.low for s in subs]) if subs else np.zeros(0), np.concatenate([s.high for s in subs]) if subs else np.zeros(0), dtype=dtype)
    raise NotImplementedError(type(space))


def flatten(space, x):
    from . import Box, Dict, Discrete, Tuple
    if isinstance(space, Box):
        return np.asarray(x, dtype=space.dtype).flatten()
    if isinstance(space, Discrete):
        out = np.zeros(space.n, dtype=space.dtype)
        out[int(x) - space.start] = 1
        return out
    if isinstance(space, Dict):
        return np.concatenate([flatten(s, x[k]) for k, s in space.spaces.items()]) if space.spaces else np.zeros(0)
    if isinstance(space, Tuple):
        return np.concatenate([flatten(s, v) for s, v in zip(space.spaces, x)])
    raise NotImplementedError(type(space))
