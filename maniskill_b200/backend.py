"""ctypes binding of libb200sim.so (include/b200sim.h) -- the object ManiSkill-side code talks to instead of
``sapien.physx.PhysxGpuSystem`` (mani_skill/envs/scene.py:40-66, 379-380, 950-986).

There is NO CPU path: constructing a :class:`World` without the compiled CUDA library or without a CUDA device
raises.  The exposed buffers are torch CUDA tensors that alias the device state, exactly like
``px.cuda_rigid_body_data.torch()`` (mani_skill/utils/structs/base.py:262-270).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from .model import B2SModelStruct, CompiledModel

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200sim.so")

BUF_RIGID, BUF_ROOT_POSE, BUF_QPOS, BUF_QVEL, BUF_QF, BUF_TARGET_QPOS, BUF_TARGET_QVEL, BUF_QACC, BUF_LINK = [1 << i for i in range(9)]
BUF_ALL = 0xFFFFFFFF
BUF_APPLY_ALL = BUF_RIGID | BUF_ROOT_POSE | BUF_QPOS | BUF_QVEL | BUF_QF | BUF_TARGET_QPOS | BUF_TARGET_QVEL


class B2SBufferTable(C.Structure):
    _fields_ = [("rigid_body_data", C.c_void_p), ("qpos", C.c_void_p), ("qvel", C.c_void_p), ("qacc", C.c_void_p),
                ("qf", C.c_void_p), ("target_qpos", C.c_void_p), ("target_qvel", C.c_void_p), ("n_rows", C.c_int32),
                ("max_dof", C.c_int32), ("contact_count", C.c_void_p), ("overflow_flag", C.c_void_p)]


class B2SCameraDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("near_", C.c_float), ("far_", C.c_float), ("mount_row", C.c_int32),
                ("local_pose", C.c_float * 7)]


class B2SVisualTable(C.Structure):
    _fields_ = [("n_visual", C.c_int32), ("type", C.c_void_p), ("row", C.c_void_p), ("pose", C.c_void_p), ("size", C.c_void_p),
                ("color", C.c_void_p), ("seg_id", C.c_void_p), ("ov_slot", C.c_void_p), ("n_ov", C.c_int32), ("ov_size", C.c_void_p),
                ("ov_pose", C.c_void_p), ("n_vert", C.c_int32), ("vert_local", C.c_void_p), ("vert_vis", C.c_void_p), ("n_tri", C.c_int32),
                ("tri_idx", C.c_void_p), ("tri_vis", C.c_void_p)]


class B2SRenderTargets(C.Structure):
    _fields_ = [("color", C.c_void_p), ("position_seg", C.c_void_p), ("rgb", C.c_void_p), ("depth", C.c_void_p), ("segmentation", C.c_void_p)]


# include/b200sim.h B2S_OUT_*: what a camera group writes per pixel
OUT_COLOR, OUT_POSSEG, OUT_RGB, OUT_DEPTH, OUT_SEG = 1, 2, 4, 8, 16
OUT_RAW = OUT_COLOR | OUT_POSSEG


class B2SJointController(C.Structure):
    _fields_ = [("n_action", C.c_int32), ("dof_action", C.c_void_p), ("dof_use_delta", C.c_void_p), ("dof_normalize", C.c_void_p),
                ("dof_low", C.c_void_p), ("dof_high", C.c_void_p)]


class B2SPickTask(C.Structure):
    _fields_ = [("tcp_row", C.c_int32), ("obj_row", C.c_int32), ("goal_row", C.c_int32), ("lfinger_row", C.c_int32), ("rfinger_row", C.c_int32),
                ("goal_thresh", C.c_float), ("min_force", C.c_float), ("max_angle_deg", C.c_float), ("static_thresh", C.c_float),
                ("n_static_dof", C.c_int32), ("max_episode_steps", C.c_int32), ("normalized_reward", C.c_int32)]


class B2SPickOutputs(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("flags", C.c_void_p), ("elapsed", C.c_void_p)]


class B2SPickReset(C.Structure):
    _fields_ = [("cube_spawn_half_size", C.c_float), ("cube_spawn_center", C.c_float * 2), ("cube_half_size", C.c_float),
                ("max_goal_height", C.c_float), ("robot_qpos_noise", C.c_float), ("n_rest", C.c_int32), ("rest_qpos", C.c_float * 16),
                ("obj_fb", C.c_int32), ("goal_fb", C.c_int32)]


class B2SChainDesc(C.Structure):
    _fields_ = [("n_elem", C.c_int32), ("origin", C.c_void_p), ("axis", C.c_void_p), ("kind", C.c_void_p), ("qpos_column", C.c_void_p),
                ("controlled", C.c_void_p), ("lambda_", C.c_float), ("alpha", C.c_float)]


class B2SPickAutoReset(C.Structure):
    _fields_ = [("rand", C.c_void_p), ("final_obs", C.c_void_p), ("done", C.c_void_p), ("ignore_terminations", C.c_int32),
                ("max_episode_steps", C.c_int32)]


_lib = None


def load_library():
    """Load libb200sim.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(b200sim has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.b2s_last_error.restype = C.c_char_p
    lib.b2s_world_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]
    lib.b2s_world_destroy.argtypes = [C.c_uint64]
    lib.b2s_world_buffers.argtypes = [C.c_uint64, C.POINTER(B2SBufferTable)]
    lib.b2s_step.argtypes = [C.c_uint64, C.c_int32, C.c_uint32, C.c_void_p]
    lib.b2s_apply.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
    lib.b2s_fetch.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
    lib.b2s_update_kinematics.argtypes = [C.c_uint64, C.c_void_p]
    lib.b2s_contact_query_create.argtypes = [C.c_uint64, C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]
    lib.b2s_contact_query_run.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.b2s_camera_group_create.argtypes = [C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(B2SRenderTargets)]
    lib.b2s_camera_group_create_outputs.argtypes = [C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64),
                                                    C.POINTER(B2SRenderTargets)]
    lib.b2s_render.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    lib.b2s_render_masked.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.b2s_masked_copy.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.b2s_pick_task_autoreset.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b2s_pick_task_create.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    lib.b2s_pick_task_step.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.b2s_pick_task_set_reset.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    lib.b2s_pick_task_step_autoreset.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.b2s_ik_create.argtypes = [C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
    lib.b2s_ik_step.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ["b2s_last_error", "b2s_version", "b2s_world_create", "b2s_world_destroy", "b2s_world_buffers", "b2s_step",
                    "b2s_apply", "b2s_fetch", "b2s_update_kinematics", "b2s_contact_query_create", "b2s_contact_query_run",
                    "b2s_camera_group_create", "b2s_camera_group_create_outputs", "b2s_render", "b2s_pick_task_create", "b2s_pick_task_step", "b2s_pick_task_set_reset",
                    "b2s_pick_task_step_autoreset", "b2s_pick_task_autoreset", "b2s_masked_copy", "b2s_render_masked", "b2s_ik_create", "b2s_ik_step"]


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can alias library-owned device memory (zero copy)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)
        self._owner = owner


def _as_tensor(ptr, shape, typestr, owner, device):
    return torch.as_tensor(_DevArray(ptr, shape, typestr, owner), device=device)


def _check(lib, code):
    if code != 0:
        raise RuntimeError(f"b200sim error {code}: {lib.b2s_last_error().decode()}")


class CameraGroup:
    """What ``render_system_group.create_camera_group(...)`` returns in the reference (mani_skill/envs/scene.py:1087-1106):
    ``take_picture()`` renders every camera of every sub-scene, ``get_picture_cuda(name)`` hands out zero-copy tensors."""

    def __init__(self, world, handle, cameras, color, posseg, rgb=None, depth=None, seg=None):
        self.world, self.handle, self.cameras = world, handle, cameras
        # raw render targets [N, P, 4] (None when the group only writes compact textures) and the compact textures
        # rgb [N, P, 3] uint8, depth [N, P] int16, segmentation [N, P] int16 (None unless requested)
        self._color, self._posseg = color, posseg
        self._rgb, self._depth, self._seg = rgb, depth, seg
        self._final = None   # same-named copies kept for finished sub-scenes (render.CameraSensors.keep_final)
        self._offsets = np.cumsum([0] + [int(c["width"]) * int(c["height"]) for c in cameras])

    def buffers(self):
        """name -> tensor of every output this group writes."""
        return {k: v for k, v in (("color", self._color), ("posseg", self._posseg), ("rgb", self._rgb), ("depth", self._depth), ("seg", self._seg))
                if v is not None}

    def texture(self, name: str, cam: int = 0, final: bool = False):
        """Compact texture of one camera: 'rgb' [N,H,W,3] uint8 | 'depth' [N,H,W,1] int16 | 'seg' [N,H,W,1] int16, or None when the group
        does not write it."""
        src = self._final if final else self.buffers()
        buf = src.get(name)
        if buf is None:
            return None
        c = self.cameras[cam]
        a, b = int(self._offsets[cam]), int(self._offsets[cam + 1])
        return buf[:, a:b].view(self.world.n_envs, int(c["height"]), int(c["width"]), 3 if name == "rgb" else 1)

    def take_picture(self):
        self.world.render(self)

    def get_picture_cuda(self, name: str, cam: int = 0, final: bool = False):
        """[N, H, W, 4] view of one camera's render target ('Color' uint8 | 'PositionSegmentation' int16); final=True: the copy kept for
        sub-scenes that an auto-reset re-rendered (maniskill_b200/render.py `keep_final`)."""
        c = self.cameras[cam]
        a, b = int(self._offsets[cam]), int(self._offsets[cam + 1])
        key = "color" if name == "Color" else "posseg"
        buf = (self._final if final else self.buffers()).get(key)
        if buf is None:
            raise RuntimeError(f"this camera group does not write the raw render target '{name}' (created with compact outputs only)")
        return buf[:, a:b].view(self.world.n_envs, int(c["height"]), int(c["width"]), 4)


# bits of the sticky overflow word (include/b200sim.h B2S_OVF_*) -> the capacity that was exceeded and how to raise it
OVERFLOW_REASONS = {
    1: "contact patches per sub-scene > max_manifolds (sim_config max_manifolds, compiled cap 24)",
    2: "contact points per sub-scene > max_contacts (sim_config max_contacts, compiled cap 64)",
    4: "constraint rows per sub-scene > the compiled row capacity (csrc/b2s_step.cuh Caps::MAXROW)",
    8: "constraint rows on an articulation > capacity",
    16: "active joint limits per sub-scene > 24",
    32: "tendon couplings per sub-scene > 2",
}


class CapacityWarning(RuntimeWarning):
    """A fixed per-sub-scene capacity dropped a contact or a constraint row (the reference's PhysX reports exceeded
    GPUMemoryConfig capacities in the same spirit, mani_skill/utils/structs/types.py:16-32)."""


ANY_BODY = -2  # include/b200sim.h B2S_ANY_BODY: second row of a contact query = every body (net impulse on the first)


class World:
    """One batched world on one GPU (the PhysxGpuSystem + all sub-scenes of the reference)."""

    def __init__(self, cm: CompiledModel, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("b200sim needs a CUDA device (B200); there is no CPU path")
        self.lib = load_library()
        self.cm = cm
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._struct = cm.struct()
        h = C.c_uint64(0)
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.b2s_world_create(C.addressof(self._struct), self.device.index, C.byref(h)))
        self.h = h
        tab = B2SBufferTable()
        _check(self.lib, self.lib.b2s_world_buffers(self.h, C.byref(tab)))
        s = cm.scalars
        self.n_envs, self.n_rows, self.n_link, self.n_fb, self.n_art = s["n_envs"], tab.n_rows, s["n_link"], s["n_fb"], s["n_art"]
        self.max_dof = max(tab.max_dof, 1)
        N = self.n_envs
        f4 = "<f4"
        self.rigid_body_data = _as_tensor(tab.rigid_body_data, (N * self.n_rows, 13), f4, self, self.device)
        shape_q = (N * max(self.n_art, 1), self.max_dof)
        self.qpos = _as_tensor(tab.qpos, shape_q, f4, self, self.device)
        self.qvel = _as_tensor(tab.qvel, shape_q, f4, self, self.device)
        self.qacc = _as_tensor(tab.qacc, shape_q, f4, self, self.device)
        self.qf = _as_tensor(tab.qf, shape_q, f4, self, self.device)
        self.target_qpos = _as_tensor(tab.target_qpos, shape_q, f4, self, self.device)
        self.target_qvel = _as_tensor(tab.target_qvel, shape_q, f4, self, self.device)
        self.contact_count = _as_tensor(tab.contact_count, (N,), "<i4", self, self.device)
        self.overflow_flag = _as_tensor(tab.overflow_flag, (1,), "<i4", self, self.device)
        self._queries = {}
        self.kernel_launches = 0

    # ------------------------------------------------------------------ stream
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _step_launches(self, substeps, fetch_mask):
        """Kernels one b2s_step launches (csrc/b2s_api.cu): kin x2, collide, manifest, rowfill, solve per substep + the fetch
        (replayed as one CUDA graph); B2S_FUSED=1 selects the single-lane kernel."""
        if os.environ.get("B2S_FUSED", "0") not in ("", "0"):
            return 1
        return 6 * int(substeps) + (1 if fetch_mask else 0)

    # ------------------------------------------------------------------ px.* entry points
    def step(self, substeps: int = 1, fetch_mask: int = 0):
        _check(self.lib, self.lib.b2s_step(self.h, substeps, fetch_mask, self._stream()))
        self.kernel_launches += self._step_launches(substeps, fetch_mask)

    def apply(self, mask: int = BUF_APPLY_ALL):
        _check(self.lib, self.lib.b2s_apply(self.h, mask, self._stream()))
        self.kernel_launches += 1

    def fetch(self, mask: int = BUF_ALL):
        _check(self.lib, self.lib.b2s_fetch(self.h, mask, self._stream()))
        self.kernel_launches += 1

    def update_kinematics(self):
        _check(self.lib, self.lib.b2s_update_kinematics(self.h, self._stream()))
        self.kernel_launches += 1

    def body_view(self):
        """[n_envs, n_rows, 13] view of rigid_body_data."""
        return self.rigid_body_data.view(self.n_envs, self.n_rows, 13)

    def create_contact_query(self, row_pairs):
        key = tuple(map(tuple, row_pairs))
        if key not in self._queries:
            rows = np.ascontiguousarray(np.asarray(row_pairs, dtype=np.int32).reshape(-1))
            q = C.c_uint64(0)
            _check(self.lib, self.lib.b2s_contact_query_create(self.h, rows.ctypes.data_as(C.c_void_p), len(row_pairs), C.byref(q)))
            out = torch.zeros((self.n_envs, len(row_pairs), 3), dtype=torch.float32, device=self.device)
            self._queries[key] = (q, out)
        return key

    def query_contact_impulses(self, key):
        q, out = self._queries[key]
        _check(self.lib, self.lib.b2s_contact_query_run(self.h, q, C.c_void_p(out.data_ptr()), self._stream()))
        self.kernel_launches += 1
        return out

    # ------------------------------------------------------------------ fused control step (pick task family)
    def create_pick_task(self, dof_action, dof_use_delta, dof_normalize, dof_low, dof_high, n_action, rows: dict, goal_thresh,
                         n_static_dof, max_episode_steps, normalized_reward=True, min_force=0.5, max_angle_deg=85.0, static_thresh=0.2):
        ctrl = B2SJointController()
        keep = [np.ascontiguousarray(np.asarray(dof_action, dtype=np.int32)), np.ascontiguousarray(np.asarray(dof_use_delta, dtype=np.int32)),
                np.ascontiguousarray(np.asarray(dof_normalize, dtype=np.int32)), np.ascontiguousarray(np.asarray(dof_low, dtype=np.float32)),
                np.ascontiguousarray(np.asarray(dof_high, dtype=np.float32))]
        ctrl.n_action = int(n_action)
        ctrl.dof_action, ctrl.dof_use_delta, ctrl.dof_normalize, ctrl.dof_low, ctrl.dof_high = [a.ctypes.data_as(C.c_void_p) for a in keep]
        task = B2SPickTask(int(rows["tcp"]), int(rows["obj"]), int(rows["goal"]), int(rows["lfinger"]), int(rows["rfinger"]), float(goal_thresh),
                           float(min_force), float(max_angle_deg), float(static_thresh), int(n_static_dof), int(max_episode_steps or 0),
                           1 if normalized_reward else 0)
        h = C.c_uint64(0)
        _check(self.lib, self.lib.b2s_pick_task_create(self.h, C.byref(ctrl), C.byref(task), C.byref(h)))
        return h

    def pick_task_step(self, handle, actions, substeps, obs, reward, flags, elapsed):
        out = B2SPickOutputs(obs.data_ptr(), reward.data_ptr(), flags.data_ptr(), elapsed.data_ptr())
        a = C.c_void_p(actions.data_ptr()) if actions is not None else None
        _check(self.lib, self.lib.b2s_pick_task_step(self.h, handle, a, int(substeps), C.byref(out), self._stream()))
        self.kernel_launches += (2 if actions is not None else 1) + self._step_launches(substeps, BUF_ALL)

    def set_pick_reset(self, handle, cube_spawn_half_size, cube_spawn_center, cube_half_size, max_goal_height, robot_qpos_noise, rest_qpos,
                       obj_fb, goal_fb):
        """Episode initialisation of the pick task for the device-side auto-reset (include/b200sim.h B2SPickReset)."""
        r = B2SPickReset()
        r.cube_spawn_half_size, r.cube_half_size, r.max_goal_height = float(cube_spawn_half_size), float(cube_half_size), float(max_goal_height)
        r.cube_spawn_center = (C.c_float * 2)(float(cube_spawn_center[0]), float(cube_spawn_center[1]))
        r.robot_qpos_noise = float(robot_qpos_noise)
        rest = [float(x) for x in rest_qpos]
        r.n_rest = len(rest)
        r.rest_qpos = (C.c_float * 16)(*(rest + [0.0] * (16 - len(rest))))
        r.obj_fb, r.goal_fb = int(obj_fb), int(goal_fb)
        _check(self.lib, self.lib.b2s_pick_task_set_reset(self.h, handle, C.byref(r)))

    def pick_task_autoreset(self, handle, obs, reward, flags, elapsed, rand, final_obs, done, ignore_terminations=False, max_episode_steps=0):
        """The auto-reset alone, after `pick_task_step` (the caller may render the finished state in between)."""
        out = B2SPickOutputs(obs.data_ptr(), reward.data_ptr(), flags.data_ptr(), elapsed.data_ptr())
        ar = B2SPickAutoReset(rand.data_ptr(), final_obs.data_ptr(), done.data_ptr(), 1 if ignore_terminations else 0, int(max_episode_steps or 0))
        _check(self.lib, self.lib.b2s_pick_task_autoreset(self.h, handle, C.byref(out), C.byref(ar), self._stream()))
        self.kernel_launches += 3

    def pick_task_step_autoreset(self, handle, actions, substeps, obs, reward, flags, elapsed, rand, final_obs, done, ignore_terminations=False,
                                 max_episode_steps=0):
        """b2s_pick_task_step + the auto-reset of finished sub-scenes, all on the device (no host sync)."""
        out = B2SPickOutputs(obs.data_ptr(), reward.data_ptr(), flags.data_ptr(), elapsed.data_ptr())
        ar = B2SPickAutoReset(rand.data_ptr(), final_obs.data_ptr(), done.data_ptr(), 1 if ignore_terminations else 0, int(max_episode_steps or 0))
        a = C.c_void_p(actions.data_ptr()) if actions is not None else None
        _check(self.lib, self.lib.b2s_pick_task_step_autoreset(self.h, handle, a, int(substeps), C.byref(out), C.byref(ar), self._stream()))
        self.kernel_launches += (2 if actions is not None else 1) + self._step_launches(substeps, BUF_ALL) + 3

    # ------------------------------------------------------------------ end-effector controllers
    def create_ik(self, origin7, axis3, kind, qpos_column, controlled, lambd=1e-4, alpha=1.0):
        """Serial chain root -> end link (include/b200sim.h B2SChainDesc) -> handle for `ik_step`."""
        o = np.ascontiguousarray(origin7, dtype=np.float32).reshape(-1, 7)
        a = np.ascontiguousarray(axis3, dtype=np.float32).reshape(-1, 3)
        k = np.ascontiguousarray(kind, dtype=np.int32)
        c = np.ascontiguousarray(qpos_column, dtype=np.int32)
        m = np.ascontiguousarray(controlled, dtype=np.uint8)
        d = B2SChainDesc(len(k), o.ctypes.data, a.ctypes.data, k.ctypes.data, c.ctypes.data, m.ctypes.data, float(lambd), float(alpha))
        h = C.c_uint64(0)
        _check(self.lib, self.lib.b2s_ik_create(self.h, C.byref(d), C.byref(h)))
        return h.value

    def ik_step(self, handle, delta_pose, qpos, qpos_stride, target):
        """One damped least-squares step: target[env] = q + alpha J^T (J J^T + lambda I)^-1 delta_pose[env] for the controlled joints."""
        _check(self.lib, self.lib.b2s_ik_step(self.h, C.c_uint64(handle), C.c_void_p(delta_pose.data_ptr()), C.c_void_p(qpos.data_ptr()), int(qpos_stride),
                                              C.c_void_p(target.data_ptr()), self._stream()))
        self.kernel_launches += 1

    # ------------------------------------------------------------------ rendering
    def create_camera_group(self, cameras, visuals, outputs: int = OUT_RAW):
        """cameras: list of dict(width, height, fx, fy, cx, cy, near, far, mount_row, local_pose7);
        visuals: dict of numpy arrays (see maniskill_b200/render.py).  Returns a CameraGroup with aliasing tensors
        ``color`` [N, P, 4] uint8 and ``position_seg`` [N, P, 4] int16 (P = pixels of all cameras of one sub-scene)."""
        n_cam = len(cameras)
        arr = (B2SCameraDesc * n_cam)()
        for i, c in enumerate(cameras):
            arr[i].width, arr[i].height = int(c["width"]), int(c["height"])
            arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = c["fx"], c["fy"], c["cx"], c["cy"]
            arr[i].near_, arr[i].far_ = c["near"], c["far"]
            arr[i].mount_row = int(c["mount_row"])
            arr[i].local_pose = (C.c_float * 7)(*[float(x) for x in c["local_pose"]])
        keep = {k: np.ascontiguousarray(v) for k, v in visuals.items() if isinstance(v, np.ndarray)}
        for k in keep:
            if keep[k].size == 0:
                keep[k] = np.zeros(1, dtype=keep[k].dtype)
        vt = B2SVisualTable()
        vt.n_visual, vt.n_ov, vt.n_vert, vt.n_tri = int(visuals["n_visual"]), int(visuals["n_ov"]), int(visuals["n_vert"]), int(visuals["n_tri"])
        for name in ("type", "row", "pose", "size", "color", "seg_id", "ov_slot", "ov_size", "ov_pose", "vert_local", "vert_vis", "tri_idx", "tri_vis"):
            setattr(vt, name, keep[name].ctypes.data_as(C.c_void_p))
        g = C.c_uint64(0)
        rt = B2SRenderTargets()
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.b2s_camera_group_create_outputs(self.h, C.cast(arr, C.c_void_p), n_cam, C.byref(vt), int(outputs), C.byref(g),
                                                                      C.byref(rt)))
        pix = sum(int(c["width"]) * int(c["height"]) for c in cameras)
        t = lambda ptr, shape, ty: _as_tensor(ptr, shape, ty, self, self.device) if ptr else None
        return CameraGroup(self, g, cameras, t(rt.color, (self.n_envs, pix, 4), "|u1"), t(rt.position_seg, (self.n_envs, pix, 4), "<i2"),
                           t(rt.rgb, (self.n_envs, pix, 3), "|u1"), t(rt.depth, (self.n_envs, pix), "<i2"), t(rt.segmentation, (self.n_envs, pix), "<i2"))

    def render(self, group, env_mask=None):
        """take_picture(); `env_mask` ([n_envs] uint8 / bool device tensor): only those sub-scenes are re-rendered."""
        if env_mask is None:
            _check(self.lib, self.lib.b2s_render(self.h, group.handle, self._stream()))
        else:
            _check(self.lib, self.lib.b2s_render_masked(self.h, group.handle, C.c_void_p(env_mask.data_ptr()), self._stream()))
        self.kernel_launches += 1

    def masked_copy(self, dst, src, env_mask):
        """dst[env] = src[env] where env_mask[env] (both [n_envs, ...] contiguous device tensors of the same layout)."""
        row_bytes = src[0].numel() * src.element_size()
        _check(self.lib, self.lib.b2s_masked_copy(self.h, C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), row_bytes,
                                                  C.c_void_p(env_mask.data_ptr()), self._stream()))
        self.kernel_launches += 1

    # ------------------------------------------------------------------ capacity overflow (sticky, device side)
    def overflow_reasons(self):
        """Reads the sticky overflow word (one device sync) -> list of messages, empty when nothing was ever dropped."""
        code = int(self.overflow_flag.item())
        return [msg for bit, msg in OVERFLOW_REASONS.items() if code & bit]

    def check_overflow(self, strict: bool = False):
        """Warn (CapacityWarning) or raise (strict) when a sub-scene exceeded a capacity since the world was created.  Called by the
        env on full resets and on close, i.e. off the per-step path: the check costs a device sync."""
        reasons = self.overflow_reasons()
        if reasons:
            msg = "b200sim dropped contacts / constraint rows: " + "; ".join(reasons)
            if strict:
                raise RuntimeError(msg)
            import warnings
            warnings.warn(msg, CapacityWarning, stacklevel=2)
        return reasons

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            torch.cuda.synchronize(self.device)
            try:
                self.check_overflow()
            except Exception:
                pass
            self.lib.b2s_world_destroy(self.h)
            self.h = C.c_uint64(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
