"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around the CPU oracle (oracle/b2s_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (maniskill_b200/) never does; it fails loudly when the CUDA library is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_D = C.POINTER(C.c_double)


def build(force=False):
    """Compile the two oracle builds (float32 / float64) with the committed Makefile."""
    # make tracks the dependencies (sources + include/b200sim.h); a no-op when everything is up to date
    subprocess.check_call(["make", "-C", _DIR, "-s"] + (["-B"] if force else []))


_libs = {}


def _lib(precision):
    if precision not in _libs:
        build()
        lib = C.CDLL(os.path.join(_DIR, f"libb2s_oracle_{precision}.so"))
        lib.b2o_create.restype = C.c_void_p
        lib.b2o_create.argtypes = [C.c_void_p]
        for name in ("b2o_destroy",):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.b2o_set_joint.argtypes = [C.c_void_p, C.c_int, _D]
        lib.b2o_get_joint.argtypes = [C.c_void_p, C.c_int, _D]
        lib.b2o_set_bodies.argtypes = [C.c_void_p, _D]
        lib.b2o_get_bodies.argtypes = [C.c_void_p, _D]
        lib.b2o_set_roots.argtypes = [C.c_void_p, _D]
        lib.b2o_get_links.argtypes = [C.c_void_p, _D]
        lib.b2o_step.argtypes = [C.c_void_p, C.c_int]
        lib.b2o_step_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.b2o_step_mt.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.b2o_contact_count.argtypes = [C.c_void_p, C.c_int]
        lib.b2o_get_contacts.argtypes = [C.c_void_p, C.c_int, _D]
        lib.b2o_pair_impulse.argtypes = [C.c_void_p, C.c_int, C.c_int, _D]
        lib.b2o_overflow.argtypes = [C.c_void_p]
        lib.b2o_collide.argtypes = [C.c_int, _D, _D, C.POINTER(C.c_float), C.c_int, C.c_int, _D, _D, C.POINTER(C.c_float), C.c_int, C.c_double, _D]
        _libs[precision] = lib
    return _libs[precision]


JOINT_FIELDS = dict(qpos=0, qvel=1, target_qpos=2, target_qvel=3, qf=4, qacc=5)


class OracleWorld:
    """CPU restatement of one batched world. State arrays are float64 numpy, env-major."""

    def __init__(self, compiled_model, precision="f32"):
        self.cm = compiled_model
        self.lib = _lib(precision)
        self._struct = compiled_model.struct()
        self.h = self.lib.b2o_create(C.addressof(self._struct))
        s = compiled_model.scalars
        self.n_envs, self.n_dof, self.n_fb, self.n_link, self.n_art = s["n_envs"], s["n_dof"], s["n_fb"], s["n_link"], s["n_art"]

    def __del__(self):
        try:
            self.lib.b2o_destroy(self.h)
        except Exception:
            pass

    def set_joint(self, name, arr):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(arr, dtype=np.float64), (self.n_envs, self.n_dof)))
        self.lib.b2o_set_joint(self.h, JOINT_FIELDS[name], a.ctypes.data_as(_D))

    def get_joint(self, name):
        a = np.zeros((self.n_envs, self.n_dof))
        self.lib.b2o_get_joint(self.h, JOINT_FIELDS[name], a.ctypes.data_as(_D))
        return a

    def set_bodies(self, arr):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(arr, dtype=np.float64), (self.n_envs, self.n_fb, 13)))
        self.lib.b2o_set_bodies(self.h, a.ctypes.data_as(_D))

    def get_bodies(self):
        a = np.zeros((self.n_envs, self.n_fb, 13))
        self.lib.b2o_get_bodies(self.h, a.ctypes.data_as(_D))
        return a

    def set_roots(self, arr):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(arr, dtype=np.float64), (self.n_envs, self.n_art, 7)))
        self.lib.b2o_set_roots(self.h, a.ctypes.data_as(_D))

    def get_links(self):
        a = np.zeros((self.n_envs, self.n_link, 13))
        self.lib.b2o_get_links(self.h, a.ctypes.data_as(_D))
        return a

    def rigid_body_data(self):
        """[n_envs, n_link+n_fb, 13] exactly like the exposed cuda_rigid_body_data rows."""
        return np.concatenate([self.get_links(), self.get_bodies()], axis=1)

    def step(self, substeps=1):
        self.lib.b2o_step(self.h, substeps)

    def step_range(self, substeps, lo, hi):
        self.lib.b2o_step_range(self.h, substeps, lo, hi)

    def step_mt(self, substeps, nthreads):
        self.lib.b2o_step_mt(self.h, substeps, nthreads)

    def contacts(self, env=0):
        n = self.lib.b2o_contact_count(self.h, env)
        a = np.zeros((max(n, 1), 12))
        self.lib.b2o_get_contacts(self.h, env, a.ctypes.data_as(_D))
        return a[:n]

    def pair_impulse(self, row_a, row_b):
        a = np.zeros((self.n_envs, 3))
        self.lib.b2o_pair_impulse(self.h, row_a, row_b, a.ctypes.data_as(_D))
        return a

    def overflow(self):
        return self.lib.b2o_overflow(self.h)


def collide(shape_a, shape_b, margin=0.04, precision="f64"):
    """Narrowphase probe. shape = dict(type, pose7, size3, verts[n,3] optional). Returns [n,7] (p, n, sep)."""
    lib = _lib(precision)

    def prep(s):
        pose = np.ascontiguousarray(np.asarray(s["pose"], dtype=np.float64))
        size = np.ascontiguousarray(np.asarray(s.get("size", [0, 0, 0]), dtype=np.float64))
        v = np.ascontiguousarray(np.asarray(s.get("verts", np.zeros((1, 3))), dtype=np.float32))
        return pose, size, v

    pa, sa, va = prep(shape_a)
    pb, sb, vb = prep(shape_b)
    out = np.zeros((4, 7))
    F = C.POINTER(C.c_float)
    n = lib.b2o_collide(shape_a["type"], pa.ctypes.data_as(_D), sa.ctypes.data_as(_D), va.ctypes.data_as(F), len(va),
                        shape_b["type"], pb.ctypes.data_as(_D), sb.ctypes.data_as(_D), vb.ctypes.data_as(F), len(vb),
                        margin, out.ctypes.data_as(_D))
    return out[:n]
