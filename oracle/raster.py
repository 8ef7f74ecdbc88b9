"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the CPU raster oracle (oracle/b2s_oracle_raster.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR, "-s"])
        _lib = C.CDLL(os.path.join(_DIR, "libb2s_oracle_raster.so"))
        _lib.b2o_render_one.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    return _lib


def render(visuals: dict, cameras: list, body: np.ndarray):
    """visuals: maniskill_b200.render.build_visual_table output; cameras: list of camera_desc dicts;
    body: [N, n_rows, 13] float32.  Returns per camera (color [N,H,W,4] uint8, posseg [N,H,W,4] int16)."""
    L = lib()
    body = np.ascontiguousarray(body, dtype=np.float32)
    N, n_rows = body.shape[0], body.shape[1]
    nv, n_ov = int(visuals["n_visual"]), int(visuals["n_ov"])
    out = []
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for cam in cameras:
        H, W = cam["height"], cam["width"]
        color = np.zeros((N, H, W, 4), dtype=np.uint8)
        posseg = np.zeros((N, H, W, 4), dtype=np.int16)
        camv = np.array([W, H, cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["near"], cam["far"], cam["mount_row"]] + list(cam["local_pose"]), dtype=np.float32)
        for e in range(N):
            pose = visuals["pose"].reshape(nv, 7).copy()
            size = visuals["size"].reshape(nv, 3).copy()
            if n_ov:
                ovs = visuals["ov_size"].reshape(N, n_ov, 3)
                ovp = visuals["ov_pose"].reshape(N, n_ov, 7)
                for v in range(nv):
                    s = visuals["ov_slot"][v]
                    if s >= 0:
                        size[v] = ovs[e, s]
                        pose[v] = ovp[e, s]
            pose = np.ascontiguousarray(pose, dtype=np.float32)
            size = np.ascontiguousarray(size, dtype=np.float32)
            L.b2o_render_one(nv, p(visuals["type"]), p(visuals["row"]), p(pose), p(size), p(visuals["color"]), p(visuals["seg_id"]),
                             int(visuals["n_vert"]), p(visuals["vert_local"]), p(visuals["vert_vis"]), int(visuals["n_tri"]), p(visuals["tri_idx"]),
                             p(visuals["tri_vis"]), n_rows, p(body[e]), p(camv),
                             p(color[e]), p(posseg[e]))
        out.append((color, posseg))
    return out
