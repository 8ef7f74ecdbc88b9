"""gymnasium.spaces: Space, Box, Discrete, MultiDiscrete, MultiBinary, Dict, Tuple, Text (the subset mani_skill builds)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Optional, Sequence

import numpy as np

from . import utils  # noqa: F401,E402  (defined at the bottom through a late import)


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]

    def sample(self, mask=None):
        raise NotImplementedError

    def contains(self, x) -> bool:
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    @property
    def is_np_flattenable(self):
        return False


class Box(Space):
    def __init__(self, low, high, shape: Optional[Sequence[int]] = None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        super().__init__(shape, dtype, seed)

    @property
    def is_np_flattenable(self):
        return True

    def is_bounded(self, manner="both"):
        below, above = bool(np.all(self.bounded_below)), bool(np.all(self.bounded_above))
        return {"both": below and above, "below": below, "above": above}[manner]

    def sample(self, mask=None):
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp = ~self.bounded_below & self.bounded_above
        low_b = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        rng = self.np_random
        sample[unbounded] = rng.normal(size=unbounded[unbounded].shape)
        sample[low_b] = rng.exponential(size=low_b[low_b].shape) + self.low[low_b]
        sample[upp] = -rng.exponential(size=upp[upp].shape) + self.high[upp]
        sample[bounded] = rng.uniform(low=self.low[bounded], high=high[bounded], size=bounded[bounded].shape)
        if self.dtype.kind in "iu":
            sample = np.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min() if self.low.size else 0}, {self.high.max() if self.high.size else 0}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)


class Discrete(Space):
    def __init__(self, n: int, seed=None, start: int = 0):
        self.n, self.start = int(n), int(start)
        super().__init__((), np.int64, seed)

    def sample(self, mask=None):
        return np.int64(self.start + self.np_random.integers(self.n))

    def contains(self, x) -> bool:
        return bool(self.start <= int(x) < self.start + self.n)

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self, mask=None):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= 0) and np.all(x < self.nvec))


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n
        super().__init__((n,) if np.isscalar(n) else tuple(n), np.int8, seed)

    def sample(self, mask=None):
        return self.np_random.integers(0, 2, size=self.shape, dtype=self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))


class Text(Space):
    def __init__(self, max_length: int, min_length: int = 1, charset=None, seed=None):
        self.max_length, self.min_length = max_length, min_length
        super().__init__(None, str, seed)

    def sample(self, mask=None):
        return "a" * self.min_length

    def contains(self, x) -> bool:
        return isinstance(x, str) and self.min_length <= len(x) <= self.max_length


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kw):
        if spaces is None:
            spaces = kw
        elif kw:
            spaces = dict(spaces, **kw)
        self.spaces = OrderedDict(spaces.items() if hasattr(spaces, "items") else spaces)
        super().__init__(None, None, seed)

    def seed(self, seed=None):
        out = super().seed(seed)
        for i, s in enumerate(self.spaces.values()):
            s.seed(None if seed is None else (seed if isinstance(seed, int) else 0) + i + 1)
        return out

    def sample(self, mask=None):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x) -> bool:
        return isinstance(x, dict) and set(x.keys()) == set(self.spaces.keys()) and all(self.spaces[k].contains(v) for k, v in x.items())

    def __getitem__(self, key):
        return self.spaces[key]

    def __setitem__(self, key, value):
        self.spaces[key] = value

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {s}" for k, s in self.spaces.items()) + ")"

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x) -> bool:
        return isinstance(x, (tuple, list)) and len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)
