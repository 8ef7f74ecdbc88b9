"""Host side of the camera sensors: render-shape table + camera descriptors for the batched rasteriser.

Mirror of the pieces of mani_skill/sensors/camera.py:126-253 (``Camera``), mani_skill/utils/structs/render_camera.py:77-182
(camera parameter tensors) and mani_skill/render/shaders.py:68-84 (the "minimal" pack's texture transforms) that the
visual observation modes use.  Render shapes are the collision primitives / convex hulls with a flat base colour (the
reference's .glb visual meshes and the floor texture are assets outside the hot path).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .model import SHAPE_BOX, SHAPE_CONVEX, CompiledModel


# corners of the unit cube (x fastest) and its twelve triangles
_UNIT_CUBE_VERTS = np.array([[(1 if c & 1 else -1), (1 if c & 2 else -1), (1 if c & 4 else -1)] for c in range(8)], dtype=np.float64)
_UNIT_CUBE_TRIS = np.array([[0, 1, 3], [0, 3, 2], [4, 5, 7], [4, 7, 6], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                            [1, 3, 7], [1, 7, 5]], dtype=np.int64)


def build_visual_table(cm: CompiledModel, n_envs: int, include_hidden: bool = False) -> Dict[str, np.ndarray]:
    vis = [v for v in cm.visuals if include_hidden or not v["hidden"]]
    hull_off = cm.arrays["hull_offset"]
    hv = cm.arrays["hull_verts"]
    hull_verts = hv.reshape(-1, 3) if hv.size % 3 == 0 and cm.scalars["n_hull"] > 0 else np.zeros((0, 3), dtype=np.float32)  # (a model without hulls keeps a 1-element placeholder)
    types, rows, poses, sizes, colors, segs, ov_slot = [], [], [], [], [], [], []
    ov_size, ov_pose = [], []
    vert_local, vert_vis, tri_idx, tri_vis = [], [], [], []
    for i, v in enumerate(vis):
        types.append(v["type"]); rows.append(v["row"]); poses.append(v["pose"]); sizes.append(v["size"])
        colors.append(v["color"]); segs.append(v["seg"])
        if v["per_env_size"] is not None or v["per_env_pose"] is not None:
            ov_slot.append(len(ov_size))
            ov_size.append(np.asarray(v["per_env_size"] if v["per_env_size"] is not None else np.tile(v["size"], (n_envs, 1)), dtype=np.float32))
            ov_pose.append(np.asarray(v["per_env_pose"] if v["per_env_pose"] is not None else np.tile(v["pose"], (n_envs, 1)), dtype=np.float32))
        else:
            ov_slot.append(-1)
        # indexed triangle geometry of the visuals the rasteriser draws as meshes: hulls (their vertices) and boxes (the corners of the
        # unit cube; the kernel scales them by the -- possibly per-env -- half extents)
        if v["type"] == SHAPE_CONVEX:
            h = v["hull"]
            verts = hull_verts[hull_off[h]:hull_off[h + 1]].astype(np.float64)
            tris = np.asarray(cm.hull_tris[h], dtype=np.int64)
        elif v["type"] == SHAPE_BOX:
            verts, tris = _UNIT_CUBE_VERTS, _UNIT_CUBE_TRIS
        else:
            continue
        tv = verts[tris]  # [n_tri, 3, 3]
        # outward winding (normal away from the solid's centre): the rasteriser culls back faces by the sign of the screen area
        nrm = np.cross(tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0])
        inward = np.einsum("ij,ij->i", nrm, tv.mean(1) - verts.mean(0)) < 0
        tris = tris.copy()
        tris[inward] = tris[inward][:, [0, 2, 1]]
        base = len(vert_vis)
        vert_local.append(verts.astype(np.float32))
        vert_vis.extend([i] * len(verts))
        tri_idx.append((tris + base).astype(np.int32))
        tri_vis.extend([i] * len(tris))
    f32 = lambda a, shape: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(shape))
    n = len(vis)
    return dict(
        n_visual=n, n_ov=len(ov_size), n_vert=len(vert_vis), n_tri=len(tri_vis),
        type=np.asarray(types, dtype=np.int32), row=np.asarray(rows, dtype=np.int32), pose=f32(poses, (-1,)), size=f32(sizes, (-1,)),
        color=f32(colors, (-1,)), seg_id=np.asarray(segs, dtype=np.int32), ov_slot=np.asarray(ov_slot, dtype=np.int32),
        ov_size=f32(np.stack(ov_size, 1) if ov_size else np.zeros(0), (-1,)), ov_pose=f32(np.stack(ov_pose, 1) if ov_pose else np.zeros(0), (-1,)),
        vert_local=f32(np.concatenate(vert_local) if vert_local else np.zeros(0), (-1,)), vert_vis=np.asarray(vert_vis, dtype=np.int32),
        tri_idx=np.ascontiguousarray(np.concatenate(tri_idx).reshape(-1) if tri_idx else np.zeros(0, dtype=np.int32), dtype=np.int32),
        tri_vis=np.asarray(tri_vis, dtype=np.int32),
    )


def camera_desc(uid, pose7, width, height, fov, near, far, mount_row=-1):
    """CameraConfig(uid, pose, width, height, fov, near, far) -> intrinsics like RenderCameraComponent.set_fovy:
    fy = (H/2) / tan(fov/2), fx = fy, principal point at the image centre."""
    fy = (height / 2.0) / np.tan(fov / 2.0)
    return dict(uid=uid, width=int(width), height=int(height), fx=float(fy), fy=float(fy), cx=width / 2.0, cy=height / 2.0,
                near=float(near), far=float(far), mount_row=int(mount_row), local_pose=[float(x) for x in pose7])


class CameraSensors:
    """All cameras of a task as one camera group (mani_skill/envs/scene.py:1087-1106) + the obs-facing accessors."""

    def __init__(self, world, cm: CompiledModel, cams: List[dict], include_hidden: bool = False, outputs: int = 3):
        """include_hidden: also draw the objects the task hides from its sensors (the human render cameras show them,
        sapien_env.py:1373-1374).  outputs: mask of backend.OUT_* -- the raw render targets (3, what `get_picture_cuda` hands out) and / or
        the compact textures an observation mode delivers (rgb 4, depth 8, segmentation 16): the kernel then writes only those."""
        self.world = world
        self.cams = cams
        self.visuals = build_visual_table(cm, world.n_envs, include_hidden=include_hidden)
        self.group = world.create_camera_group(cams, self.visuals, outputs) if outputs != 3 else world.create_camera_group(cams, self.visuals)

    def capture(self, env_mask=None):
        """take_picture(); env_mask ([N] bool / uint8 on the device): only those sub-scenes are rendered again."""
        if env_mask is None:
            self.group.take_picture()
        else:
            self.world.render(self.group, env_mask)

    def keep_final(self, env_mask):
        """Copies the current pictures of the masked sub-scenes into the `final` render targets (what `final_observation` shows after an
        auto-reset re-rendered them); rows of other sub-scenes keep whatever they held."""
        g = self.group
        if g._final is None:
            g._final = {k: torch.zeros_like(v) for k, v in g.buffers().items()}
        for k, v in g.buffers().items():
            self.world.masked_copy(g._final[k], v, env_mask)

    def get_obs(self, rgb=True, depth=True, segmentation=True, position=False, final=False):
        """sensor_data[uid] = {rgb [N,H,W,3] uint8, depth [N,H,W,1] int16 (mm), segmentation [N,H,W,1] int16}
        (texture_transforms of the minimal shader pack, mani_skill/render/shaders.py:74-83).  final=True reads the targets `keep_final` filled."""
        out = {}
        for i, c in enumerate(self.cams):
            d = {}
            g = self.group
            tex = lambda name: g.texture(name, i, final=final)   # the compact texture when the group writes it
            ps = None
            if position or (depth and tex("depth") is None) or (segmentation and tex("seg") is None):
                ps = g.get_picture_cuda("PositionSegmentation", i, final=final)
            if rgb:
                d["rgb"] = tex("rgb") if tex("rgb") is not None else g.get_picture_cuda("Color", i, final=final)[..., :3]
            if depth:
                d["depth"] = tex("depth") if tex("depth") is not None else -ps[..., 2:3]  # strided elementwise negation; a slice, not a gather
            if position:
                d["position"] = ps[..., :3]
            if segmentation:
                d["segmentation"] = tex("seg") if tex("seg") is not None else ps[..., 3:4]  # a view of the render target
            out[c["uid"]] = d
        return out

    def get_params(self, body_view: torch.Tensor):
        """sensor_param[uid] = extrinsic_cv [N,3,4], cam2world_gl [N,4,4], intrinsic_cv [N,3,3]
        (mani_skill/utils/structs/render_camera.py:77-155)."""
        from .structs import Pose
        dev = body_view.device
        N = body_view.shape[0]
        cache = self.__dict__.setdefault("_param_cache", {})
        out = {}
        for c in self.cams:
            uid = c["uid"]
            if uid not in cache:  # per-camera constants, and the whole entry for cameras that are not mounted on a moving body
                cache[uid] = dict(
                    local=Pose.create(torch.tensor([c["local_pose"]], dtype=torch.float32, device=dev).expand(N, 7)),
                    # OpenGL camera axes (x right, y up, z back) / OpenCV axes (x right, y down, z fwd) in the sapien camera frame
                    gl=torch.tensor([[0, 0, -1, 0], [-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32, device=dev),
                    cv=torch.tensor([[0, 0, 1, 0], [-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32, device=dev),
                    K=torch.tensor([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]], dtype=torch.float32, device=dev)[None].expand(N, 3, 3),
                    static=None)
            k = cache[uid]
            if k["static"] is not None:
                out[uid] = dict(k["static"])
                continue
            pose = Pose(body_view[:, c["mount_row"], :7]) * k["local"] if c["mount_row"] >= 0 else k["local"]
            T = pose.to_transformation_matrix()  # sapien camera frame (x fwd, y left, z up) -> world
            cam2world_gl = T @ k["gl"]
            cam2world_cv = T @ k["cv"]
            # inverse of a rigid transform: [R^T | -R^T t]
            Rt = cam2world_cv[:, :3, :3].transpose(1, 2)
            extrinsic_cv = torch.cat([Rt, -(Rt @ cam2world_cv[:, :3, 3:4])], dim=2)
            entry = dict(extrinsic_cv=extrinsic_cv, cam2world_gl=cam2world_gl, intrinsic_cv=k["K"])
            if c["mount_row"] < 0:
                k["static"] = entry
            out[uid] = dict(entry)
        return out
