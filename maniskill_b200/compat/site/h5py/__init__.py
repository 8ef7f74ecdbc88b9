"""Stand-in for `h5py` where it is not installed: the subset of its API the reference's trajectory code uses, over an in-memory tree
that is written as ONE zip archive of `.npy` members at `close()`.

Users: `RecordEpisode` (mani_skill/utils/wrappers/record.py:271, 574-705, 820-822: `File(path, "w")`, `create_group(name, track_order=True)`,
`create_dataset(name, data=, dtype=, compression=, compression_opts=)`, `del f[k]`, `f[new] = f[old]`, `len(f)`, `keys()`, `.filename`, `close()`),
the replay tool (mani_skill/trajectory/replay_trajectory.py:397, 276, 301: `File(path, "r")`, `f["traj_0"]["actions"][:]`, `k in f`) and
mani_skill/trajectory/utils/__init__.py:11-24 (`isinstance(x, h5py.Group)`, `x.keys()`, `x[k][:]`).

THE FILE IS NOT HDF5: member names are the dataset paths ("traj_0/env_states/actors/cube.npy"), attributes live in "__attrs__.json".  A file
written here is read back here and by `maniskill_b200.trajectory.load_trajectories`; the real h5py, when installed, shadows this package
(maniskill_b200.compat.install puts compat/site LAST on sys.path) and writes real HDF5.
"""
from __future__ import annotations

import io
import json
import os
import zipfile

import numpy as np

__version__ = "0.0-b200sim-standin"
_ATTRS = "__attrs__.json"


class Dataset:
    def __init__(self, name, data):
        self.name = name
        self._data = data
        self.attrs = {}

    shape = property(lambda self: self._data.shape)
    dtype = property(lambda self: self._data.dtype)
    ndim = property(lambda self: self._data.ndim)
    size = property(lambda self: self._data.size)

    def __getitem__(self, idx):
        out = self._data[idx]
        return out.copy() if isinstance(out, np.ndarray) else out

    def __setitem__(self, idx, value):
        self._data[idx] = value

    def __len__(self):
        return len(self._data)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __repr__(self):
        return f'<stand-in HDF5 dataset "{self.name}": shape {self.shape}, type "{self.dtype}">'


class Group:
    def __init__(self, name="/"):
        self.name = name
        self._items = {}
        self.attrs = {}

    def _child_name(self, key):
        return (self.name.rstrip("/") + "/" + key)

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node._items:
                if not create:
                    raise KeyError(f"Unable to open object (component not found: {part!r})")
                node._items[part] = Group(node._child_name(part))
            node = node._items[part]
            if not isinstance(node, Group) and create:
                raise ValueError(f"{part!r} is a dataset, not a group")
        return node

    def create_group(self, name, track_order=None):
        parent, _, leaf = name.strip("/").rpartition("/")
        node = self._walk(parent, create=True)
        if leaf in node._items:
            raise ValueError(f"Unable to create group (name already exists: {leaf!r})")
        g = node._items[leaf] = Group(node._child_name(leaf))
        return g

    def require_group(self, name):
        return self._walk(name, create=True)

    def create_dataset(self, name, shape=None, dtype=None, data=None, compression=None, compression_opts=None, **kwds):
        parent, _, leaf = name.strip("/").rpartition("/")
        node = self._walk(parent, create=True)
        if leaf in node._items:
            raise ValueError(f"Unable to create dataset (name already exists: {leaf!r})")
        if data is None:
            arr = np.zeros(shape if shape is not None else (), dtype=dtype or np.float32)
        else:
            arr = np.array(data, dtype=dtype, copy=True)
            if shape is not None:
                arr = arr.reshape(shape)
        d = node._items[leaf] = Dataset(node._child_name(leaf), arr)
        return d

    def __getitem__(self, key):
        node = self
        for part in [p for p in key.split("/") if p]:
            if not isinstance(node, Group) or part not in node._items:
                raise KeyError(f"Unable to open object (object {key!r} doesn't exist)")
            node = node._items[part]
        return node

    def __setitem__(self, key, value):
        parent, _, leaf = key.strip("/").rpartition("/")
        node = self._walk(parent, create=True)
        if isinstance(value, (Group, Dataset)):   # a hard link: the same object under a second name
            node._items[leaf] = value
        else:
            node._items[leaf] = Dataset(node._child_name(leaf), np.array(value, copy=True))

    def __delitem__(self, key):
        parent, _, leaf = key.strip("/").rpartition("/")
        del self._walk(parent)._items[leaf]

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __len__(self):
        return len(self._items)

    def __iter__(self):
        return iter(list(self._items))

    def keys(self):
        return self._items.keys()

    def values(self):
        return self._items.values()

    def items(self):
        return self._items.items()

    def get(self, key, default=None):
        return self[key] if key in self else default

    def visititems(self, fn, _prefix=""):
        for k, v in self._items.items():
            r = fn(_prefix + k, v)
            if r is not None:
                return r
            if isinstance(v, Group):
                r = v.visititems(fn, _prefix + k + "/")
                if r is not None:
                    return r

    def __repr__(self):
        return f'<stand-in HDF5 group "{self.name}" ({len(self)} members)>'


class File(Group):
    def __init__(self, name, mode="r", **kwds):
        super().__init__("/")
        self.filename = os.fspath(name)
        self.mode = "r" if mode == "r" else "r+"
        self._open = True
        exists = os.path.exists(self.filename)
        if mode in ("r", "r+") and not exists:
            raise FileNotFoundError(f"Unable to open file (unable to open file: name = '{self.filename}')")
        if mode in ("w-", "x") and exists:
            raise FileExistsError(self.filename)
        if mode in ("r", "r+", "a") and exists:
            self._load()
        elif mode in ("w", "w-", "x", "a"):
            self.flush()   # h5py creates the file when it is opened

    def _load(self):
        with zipfile.ZipFile(self.filename, "r") as z:
            names = z.namelist()
            for n in names:
                if n.endswith(".npy"):
                    path = n[:-4]
                    parent, _, leaf = path.rpartition("/")
                    node = self._walk(parent, create=True)
                    node._items[leaf] = Dataset("/" + path, np.load(io.BytesIO(z.read(n)), allow_pickle=False))
            if _ATTRS in names:
                meta = json.loads(z.read(_ATTRS).decode())
                for g in meta.get("groups", []):
                    self._walk(g, create=True)
                for path, attrs in meta.get("attrs", {}).items():
                    (self[path] if path else self).attrs.update(attrs)

    def flush(self):
        if self.mode == "r":
            return
        groups, attrs = [], {}
        with zipfile.ZipFile(self.filename + ".tmp", "w", zipfile.ZIP_DEFLATED) as z:
            def put(path, node):
                if isinstance(node, Dataset):
                    buf = io.BytesIO()
                    np.save(buf, node._data, allow_pickle=False)
                    z.writestr(path + ".npy", buf.getvalue())
                else:
                    groups.append(path)
                if node.attrs:
                    attrs[path] = {k: (v.tolist() if isinstance(v, (np.ndarray, np.generic)) else v) for k, v in node.attrs.items()}
            self.visititems(put)
            if self.attrs:
                attrs[""] = dict(self.attrs)
            z.writestr(_ATTRS, json.dumps({"groups": groups, "attrs": attrs}))
        os.replace(self.filename + ".tmp", self.filename)

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __bool__(self):
        return self._open


def is_hdf5(path):
    return False
