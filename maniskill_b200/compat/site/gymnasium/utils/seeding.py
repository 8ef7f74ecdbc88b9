import numpy as np


def np_random(seed=None):
    rng = np.random.default_rng(seed)
    return rng, seed
