"""sapien.math: shortest_rotation (scene.py:646)."""
import numpy as np


def shortest_rotation(source, target):
    """Quaternion (wxyz) of the shortest rotation taking direction `source` to direction `target`."""
    a = np.asarray(source, dtype=np.float64)
    b = np.asarray(target, dtype=np.float64)
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    d = float(np.dot(a, b))
    if d > 1 - 1e-12:
        return np.array([1.0, 0, 0, 0], dtype=np.float32)
    if d < -1 + 1e-12:
        axis = np.cross(a, [1.0, 0, 0])
        if np.linalg.norm(axis) < 1e-6:
            axis = np.cross(a, [0, 1.0, 0])
        axis /= np.linalg.norm(axis)
        return np.array([0.0, *axis], dtype=np.float32)
    axis = np.cross(a, b)
    q = np.array([1.0 + d, *axis])
    return (q / np.linalg.norm(q)).astype(np.float32)
