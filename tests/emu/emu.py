"""TEST TOOL ONLY: host emulation of the CUDA per-env code (see b2s_emu.cpp). Mirrors maniskill_b200.backend.World's
buffer interface with numpy arrays so the same checks run on CPU (here) and on the GPU (through the C-ABI)."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
BUF_RIGID, BUF_ROOT_POSE, BUF_QPOS, BUF_QVEL, BUF_QF, BUF_TARGET_QPOS, BUF_TARGET_QVEL, BUF_QACC, BUF_LINK = [1 << i for i in range(9)]
BUF_ALL = 0xFFFFFFFF


def _build():
    so = os.path.join(_DIR, "libb2s_emu.so")
    srcs = [os.path.join(_DIR, "b2s_emu.cpp")] + [os.path.join(_DIR, "../../maniskill_b200/csrc", f) for f in
                                                   ("b2s_math.cuh", "b2s_collide.cuh", "b2s_step.cuh", "b2s_solve.cuh", "b2s_pipe.cuh", "b2s_world.inl",
                                                    "../../include/b200sim.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-x", "c++", "-o", so, srcs[0]])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build())
        _lib.emu_create.restype = C.c_void_p
        _lib.emu_create.argtypes = [C.c_void_p]
        _lib.emu_destroy.argtypes = [C.c_void_p]
        _lib.emu_step.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        _lib.emu_step_pipe.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        _lib.emu_apply.argtypes = [C.c_void_p, C.c_uint]
        _lib.emu_fetch.argtypes = [C.c_void_p, C.c_uint]
        _lib.emu_buffer.restype = C.POINTER(C.c_float)
        _lib.emu_buffer.argtypes = [C.c_void_p, C.c_int]
        _lib.emu_man_count.restype = C.POINTER(C.c_int)
        _lib.emu_man_count.argtypes = [C.c_void_p]
        _lib.emu_overflow.argtypes = [C.c_void_p]
    return _lib


class EmuWorld:
    def __init__(self, cm):
        self.cm = cm
        self._struct = cm.struct()
        self.h = lib().emu_create(C.addressof(self._struct))
        if not self.h:   # the CUDA library reports the same condition through b2s_last_error(); without a handle nothing may be destroyed
            raise RuntimeError("emu_create failed: the model exceeds the capacities the emulation is compiled for (see stderr)")
        s = cm.scalars
        N, na, md = s["n_envs"], s["n_art"], max(s["max_dof_per_art"], 1)
        self.n_envs, self.n_rows, self.n_link = N, cm.n_rows, s["n_link"]

        def view(which, shape):
            p = lib().emu_buffer(self.h, which)
            return np.ctypeslib.as_array(p, shape=shape)

        self.rigid_body_data = view(0, (N, self.n_rows, 13))
        self.qpos, self.qvel, self.qacc, self.qf, self.target_qpos, self.target_qvel = [view(i, (N * na, md)) for i in range(1, 7)]
        self.man = view(7, (24 * 8, N))
        self.man_count = np.ctypeslib.as_array(lib().emu_man_count(self.h), shape=(N,))
        lib().emu_fetch(self.h, BUF_ALL)

    # True: the pipelined substep (kin -> collide -> manifest -> rowfill -> solve) the CUDA library runs; False: the single-lane
    # fused substep (B2S_FUSED=1 in the library)
    split = True

    def step(self, substeps=1, fetch_mask=BUF_ALL):
        (lib().emu_step_pipe if self.split else lib().emu_step)(self.h, substeps, fetch_mask)

    def apply(self, mask=BUF_ALL & ~BUF_LINK):
        lib().emu_apply(self.h, mask)

    def fetch(self, mask=BUF_ALL):
        lib().emu_fetch(self.h, mask)

    def pair_impulse(self, row_a, row_b):
        out = np.zeros((self.n_envs, 3), dtype=np.float32)
        for e in range(self.n_envs):
            for m in range(self.man_count[e]):
                ra, rb = int(self.man[m * 8, e]), int(self.man[m * 8 + 1, e])
                imp = self.man[m * 8 + 2:m * 8 + 5, e]
                any_body = row_b == -2  # B2S_ANY_BODY
                if ra == row_a and (any_body or rb == row_b):
                    out[e] += imp
                elif rb == row_a and (any_body or ra == row_b):
                    out[e] -= imp
        return out

    def overflow(self):
        return lib().emu_overflow(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().emu_destroy(self.h)
        except Exception:
            pass
