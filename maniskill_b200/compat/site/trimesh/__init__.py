"""Import-level stand-in for `trimesh` (the reference imports it at module level in its struct files; the hot path never builds a mesh).
`Trimesh` holds vertices / faces with the few derived quantities the reference reads (bounds, bounding_box, apply_transform, copy)."""
import numpy as np

from . import creation, primitives  # noqa: F401


class Trimesh:
    def __init__(self, vertices=None, faces=None, **kw):
        self.vertices = np.zeros((0, 3)) if vertices is None else np.asarray(vertices, dtype=np.float64)
        self.faces = np.zeros((0, 3), dtype=np.int64) if faces is None else np.asarray(faces, dtype=np.int64)

    @property
    def bounds(self):
        return np.stack([self.vertices.min(0), self.vertices.max(0)]) if len(self.vertices) else None

    @property
    def extents(self):
        b = self.bounds
        return None if b is None else b[1] - b[0]

    @property
    def bounding_box(self):
        b = self.bounds
        box = creation.box(extents=b[1] - b[0])
        box.vertices = box.vertices + (b[0] + b[1]) / 2
        return box

    @property
    def centroid(self):
        return self.vertices.mean(0) if len(self.vertices) else np.zeros(3)

    @property
    def volume(self):
        if not len(self.faces):
            return 0.0
        a, b, c = (self.vertices[self.faces[:, k]] for k in range(3))
        return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)

    @property
    def center_mass(self):
        """Volume centroid of a closed mesh (signed tetrahedra against the origin); the centre of the bounds for an open / empty one."""
        if len(self.faces):
            a, b, c = (self.vertices[self.faces[:, k]] for k in range(3))
            vol = np.einsum("ij,ij->i", a, np.cross(b, c)) / 6.0
            if abs(vol.sum()) > 1e-12:
                return ((a + b + c) / 4.0 * vol[:, None]).sum(0) / vol.sum()
        b = self.bounds
        return np.zeros(3) if b is None else (b[0] + b[1]) / 2

    def apply_transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        self.vertices = self.vertices @ T[:3, :3].T + T[:3, 3]
        return self

    def copy(self):
        return Trimesh(self.vertices.copy(), self.faces.copy())

    def sample(self, count):
        idx = np.random.randint(0, len(self.vertices), size=count)
        return self.vertices[idx]


class Scene:
    def __init__(self, geometry=None):
        self.geometry = geometry or {}


def load(*a, **kw):
    raise NotImplementedError("trimesh.load is not available in this stand-in (mesh files are read by maniskill_b200.meshio)")


class util:
    @staticmethod
    def concatenate(meshes):
        meshes = [m for m in meshes if m is not None]
        if not meshes:
            return Trimesh()
        off, vs, fs = 0, [], []
        for m in meshes:
            vs.append(m.vertices)
            fs.append(m.faces + off)
            off += len(m.vertices)
        return Trimesh(np.concatenate(vs), np.concatenate(fs))
