import numpy as np


def box(extents=(1, 1, 1), transform=None, **kw):
    from . import Trimesh
    h = np.asarray(extents, dtype=np.float64) / 2
    v = np.array([[(1 if c & 1 else -1), (1 if c & 2 else -1), (1 if c & 4 else -1)] for c in range(8)], dtype=np.float64) * h
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 7, 5], [4, 6, 7], [0, 5, 1], [0, 4, 5], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 7, 3], [1, 5, 7]])
    m = Trimesh(v, f)
    return m.apply_transform(transform) if transform is not None else m


def icosphere(subdivisions=3, radius=1.0, **kw):
    from . import Trimesh
    n = 12
    th, ph = np.meshgrid(np.linspace(0, np.pi, n), np.linspace(0, 2 * np.pi, 2 * n, endpoint=False), indexing="ij")
    v = radius * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], -1).reshape(-1, 3)
    return Trimesh(v, np.zeros((0, 3), dtype=np.int64))


def cylinder(radius=1.0, height=1.0, sections=24, transform=None, **kw):
    from . import Trimesh
    a = 2 * np.pi * np.arange(sections) / sections
    ring = np.stack([radius * np.cos(a), radius * np.sin(a)], 1)
    v = np.concatenate([np.concatenate([ring, np.full((sections, 1), z)], 1) for z in (-height / 2, height / 2)])
    m = Trimesh(v, np.zeros((0, 3), dtype=np.int64))
    return m.apply_transform(transform) if transform is not None else m


def capsule(height=1.0, radius=1.0, count=(12, 12), transform=None, **kw):
    m = cylinder(radius, height + 2 * radius, transform=transform)
    return m
