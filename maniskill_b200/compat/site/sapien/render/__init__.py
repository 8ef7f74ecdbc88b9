"""`sapien.render` on the b200sim rasteriser.

Render bodies / shapes / materials / lights are recorders; `RenderSystemGroup.create_camera_group(cameras, texture_names)` (one
`RenderCameraComponent` per sub-scene, mani_skill/envs/scene.py:1087-1106) creates ONE camera group of the batched rasteriser
(include/b200sim.h b2s_camera_group_create) over the visual table compiled at `gpu_init()`, and `take_picture()` /
`get_picture_cuda(name).torch()` are the C-ABI render call and the zero-copy `[N, H, W, 4]` views of its render targets
(`Color` uint8 -- float32 in [0, 1] under the "default" shader pack's picture format -- and `PositionSegmentation` int16).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .. import Component, Pose

_SHADER_DIR = {"camera": "minimal", "viewer": "default"}
_PICTURE_FORMAT = {"Color": "r8g8b8a8unorm", "ColorRaw": "r8g8b8a8unorm", "PositionSegmentation": "r16g16b16a16sint"}
_RT = dict(samples_per_pixel=32, path_depth=8, denoiser="none")


def set_camera_shader_dir(name):
    _SHADER_DIR["camera"] = str(name)


def get_camera_shader_dir():
    return _SHADER_DIR["camera"]


def set_viewer_shader_dir(name):
    _SHADER_DIR["viewer"] = str(name)


def get_viewer_shader_dir():
    return _SHADER_DIR["viewer"]


def set_picture_format(name, fmt):
    _PICTURE_FORMAT[str(name)] = str(fmt)


def set_ray_tracing_samples_per_pixel(n):
    _RT["samples_per_pixel"] = int(n)


def set_ray_tracing_path_depth(n):
    _RT["path_depth"] = int(n)


def set_ray_tracing_denoiser(name):
    _RT["denoiser"] = str(name)


def set_log_level(level):
    pass


def get_device_summary():
    return "b200sim rasteriser (CUDA)"


# ------------------------------------------------------------------------------------------------ materials, textures
class RenderTexture2D:
    def __init__(self, filename: str = None, mipmap_levels: int = 1, **kw):
        self.filename, self.mipmap_levels = filename, mipmap_levels

    def mean_color(self):
        """Average colour of the image (the rasteriser shades flat base colours)."""
        try:
            import cv2
            img = cv2.imread(self.filename, cv2.IMREAD_COLOR)
            b, g, r = (img.reshape(-1, 3).mean(0) / 255.0).tolist()
            return [r, g, b, 1.0]
        except Exception:
            return None


RenderTexture = RenderTexture2D


class RenderCubemap:
    def __init__(self, *a, **kw):
        self.args = a


class RenderMaterial:
    def __init__(self, emission=(0, 0, 0, 1), base_color=(1, 1, 1, 1), specular=0.0, roughness=1.0, metallic=0.0, transmission=0.0, ior=1.45,
                 transmission_roughness=0.0):
        self.emission = list(emission)
        self.base_color = [float(c) for c in base_color]
        self.specular, self.roughness, self.metallic = float(specular), float(roughness), float(metallic)
        self.transmission, self.ior, self.transmission_roughness = float(transmission), float(ior), float(transmission_roughness)
        self.base_color_texture = None
        self.diffuse_texture = self.normal_texture = self.roughness_texture = self.metallic_texture = self.emission_texture = self.transmission_texture = None

    def set_base_color(self, c):
        self.base_color = [float(x) for x in c]

    def get_base_color(self):
        return self.base_color

    def set_base_color_texture(self, t):
        self.base_color_texture = t

    set_diffuse_texture = set_base_color_texture

    def set_normal_texture(self, t):
        self.normal_texture = t

    def set_roughness_texture(self, t):
        self.roughness_texture = t

    def set_metallic_texture(self, t):
        self.metallic_texture = t

    def set_emission_texture(self, t):
        self.emission_texture = t

    def set_transmission_texture(self, t):
        self.transmission_texture = t

    def set_roughness(self, v):
        self.roughness = float(v)

    def set_metallic(self, v):
        self.metallic = float(v)

    def set_specular(self, v):
        self.specular = float(v)

    def set_emission(self, v):
        self.emission = list(v)

    def set_transmission(self, v):
        self.transmission = float(v)

    def set_ior(self, v):
        self.ior = float(v)

    def effective_color(self):
        if self.base_color_texture is not None:
            c = self.base_color_texture.mean_color()
            if c is not None:
                return c
        c = list(self.base_color) + [1.0] * (4 - len(self.base_color))
        return c[:4]


# ------------------------------------------------------------------------------------------------ render shapes
class RenderShape:
    kind = "none"

    def __init__(self, material: Optional[RenderMaterial] = None):
        self.material = material if material is not None else RenderMaterial()
        self.local_pose = Pose()
        self.name = ""
        self.gpu_pose_batch_index = -1
        self.per_scene_id = 0
        self.parts = []

    def set_gpu_pose_batch_index(self, i):
        self.gpu_pose_batch_index = int(i)

    def get_gpu_pose_batch_index(self):
        return self.gpu_pose_batch_index

    def get_local_pose(self):
        return self.local_pose

    def set_local_pose(self, p):
        self.local_pose = p

    def get_material(self):
        return self.material

    def get_name(self):
        return self.name

    def set_name(self, n):
        self.name = n

    def get_parts(self):
        return self.parts


class RenderShapePlane(RenderShape):
    kind = "plane"

    def __init__(self, scale=(1, 1, 1), material=None):
        super().__init__(material)
        self.scale = np.asarray(scale, dtype=np.float32).reshape(3)


class RenderShapeBox(RenderShape):
    kind = "box"

    def __init__(self, half_size, material=None):
        super().__init__(material)
        self.half_size = np.asarray(half_size, dtype=np.float32).reshape(3)


class RenderShapeSphere(RenderShape):
    kind = "sphere"

    def __init__(self, radius, material=None):
        super().__init__(material)
        self.radius = float(radius)


class RenderShapeCapsule(RenderShape):
    kind = "capsule"

    def __init__(self, radius, half_length, material=None):
        super().__init__(material)
        self.radius, self.half_length = float(radius), float(half_length)


class RenderShapeCylinder(RenderShapeCapsule):
    kind = "cylinder"


class RenderShapeTriangleMeshPart:
    def __init__(self, vertices, triangles, material):
        self.vertices, self.triangles, self.material = vertices, triangles, material

    def get_vertices(self):
        return np.asarray(self.vertices, dtype=np.float32)

    def get_triangles(self):
        return np.asarray(self.triangles, dtype=np.uint32)


class RenderShapeTriangleMesh(RenderShape):
    """A triangle mesh from a file (every primitive becomes a part with its own base colour) or from arrays."""
    kind = "mesh"
    _scaled: dict = {}

    def __init__(self, filename: str = None, scale=(1, 1, 1), material=None, vertices=None, triangles=None, normals=None, uvs=None):
        super().__init__(material)
        self.filename = filename
        self.scale = np.asarray(scale, dtype=np.float32).reshape(3)
        if filename is not None:
            from maniskill_b200 import meshio
            # every sub-scene attaches the same files: the scaled vertex arrays are shared (read-only) instead of copied N times
            key = (filename, tuple(float(x) for x in self.scale))
            scaled = RenderShapeTriangleMesh._scaled.get(key)
            if scaled is None:
                scaled = []
                for v, f, color in meshio.load_mesh_parts(filename):
                    sv = np.asarray(v, dtype=np.float64) * self.scale.astype(np.float64)
                    sv.setflags(write=False)
                    scaled.append((sv, f, color))
                RenderShapeTriangleMesh._scaled[key] = scaled
            for sv, f, color in scaled:
                mat = material if material is not None else RenderMaterial(base_color=color)
                self.parts.append(RenderShapeTriangleMeshPart(sv, f, mat))
        else:
            self.parts.append(RenderShapeTriangleMeshPart(np.asarray(vertices, dtype=np.float64) * self.scale.astype(np.float64), np.asarray(triangles), self.material))

    def get_scale(self):
        return self.scale


# ------------------------------------------------------------------------------------------------ components
class RenderBodyComponent(Component):
    def __init__(self):
        super().__init__()
        self.render_shapes: List[RenderShape] = []
        self.visibility = 1.0
        self.shading_mode = 0
        self.is_render_id_disabled = False

    def attach(self, shape: RenderShape):
        self.render_shapes.append(shape)
        return self

    def get_render_shapes(self):
        return self.render_shapes

    def set_visibility(self, v):
        self.visibility = float(v)

    def get_visibility(self):
        return self.visibility

    def set_property(self, name, value):
        pass

    def disable_render_id(self):
        self.is_render_id_disabled = True

    def enable_render_id(self):
        self.is_render_id_disabled = False

    def compute_global_aabb_tight(self):
        from maniskill_b200.compat.compile import render_shape_world_points
        pose = self.entity.pose if self.entity is not None else Pose()
        pts = np.concatenate([render_shape_world_points(s, pose) for s in self.render_shapes] or [np.zeros((1, 3))])
        return np.stack([pts.min(0), pts.max(0)])

    def _on_add_to_scene(self, scene):
        if scene.render_system is not None:
            scene.render_system.render_bodies.append(self)

    def _on_remove_from_scene(self, scene):
        if scene.render_system is not None and self in scene.render_system.render_bodies:
            scene.render_system.render_bodies.remove(self)


class _RenderLight(Component):
    def __init__(self):
        super().__init__()
        self.color = [1.0, 1.0, 1.0]
        self.shadow = False
        self.local_pose = Pose()
        self.shadow_near, self.shadow_far, self.shadow_map_size, self.shadow_half_size = 0.1, 10.0, 2048, 10.0
        self.inner_fov = self.outer_fov = 0.0
        self.direction = [0.0, 0.0, -1.0]

    def set_color(self, c):
        self.color = list(c)

    def set_shadow_parameters(self, *a, **kw):
        pass

    def set_local_pose(self, p):
        self.local_pose = p

    def set_shape(self, *a, **kw):
        pass

    def set_fov(self, *a, **kw):
        pass

    def _on_add_to_scene(self, scene):
        if scene.render_system is not None:
            scene.render_system.lights.append(self)


class RenderPointLightComponent(_RenderLight):
    pass


class RenderDirectionalLightComponent(_RenderLight):
    pass


class RenderSpotLightComponent(_RenderLight):
    pass


class RenderTexturedLightComponent(_RenderLight):
    pass


class RenderParallelogramLightComponent(_RenderLight):
    pass


class RenderCameraComponent(Component):
    """Pinhole camera (scene.py:247-297): width x height, fovy / intrinsics, near / far, a local pose in its entity (mount) frame."""

    def __init__(self, width: int, height: int, shader_dir: str = ""):
        super().__init__()
        self.width, self.height = int(width), int(height)
        self.local_pose = Pose()
        self.near, self.far = 0.01, 10.0
        self._fx = self._fy = 0.5 * self.height / np.tan(0.5 * np.deg2rad(35.0))
        self._cx, self._cy = self.width / 2.0, self.height / 2.0
        self.skew = 0.0
        self.gpu_pose_batch_index = -1
        self._group = None

    # ---- parameters
    def set_fovy(self, fovy, compute_x=True):
        self._fy = 0.5 * self.height / np.tan(0.5 * float(fovy))
        if compute_x:
            self._fx = self._fy

    def set_fovx(self, fovx, compute_y=True):
        self._fx = 0.5 * self.width / np.tan(0.5 * float(fovx))
        if compute_y:
            self._fy = self._fx

    fovy = property(lambda self: float(2 * np.arctan(0.5 * self.height / self._fy)), lambda self, v: self.set_fovy(v))
    fovx = property(lambda self: float(2 * np.arctan(0.5 * self.width / self._fx)), lambda self, v: self.set_fovx(v))
    fx = property(lambda self: float(self._fx))
    fy = property(lambda self: float(self._fy))
    cx = property(lambda self: float(self._cx))
    cy = property(lambda self: float(self._cy))

    def set_focal_lengths(self, fx, fy):
        self._fx, self._fy = float(fx), float(fy)

    def set_principal_point(self, cx, cy):
        self._cx, self._cy = float(cx), float(cy)

    def set_skew(self, s):
        self.skew = float(s)

    def set_perspective_parameters(self, near, far, fx, fy, cx, cy, skew):
        self.near, self.far, self._fx, self._fy, self._cx, self._cy, self.skew = float(near), float(far), float(fx), float(fy), float(cx), float(cy), float(skew)

    def set_near(self, v):
        self.near = float(v)

    def set_far(self, v):
        self.far = float(v)

    def get_near(self):
        return self.near

    def get_far(self):
        return self.far

    def get_width(self):
        return self.width

    def get_height(self):
        return self.height

    def set_local_pose(self, p):
        self.local_pose = p

    def get_local_pose(self):
        return self.local_pose

    def set_gpu_pose_batch_index(self, i):
        self.gpu_pose_batch_index = int(i)

    def get_gpu_pose_batch_index(self):
        return self.gpu_pose_batch_index

    def set_property(self, name, value):
        pass

    def set_texture(self, name, tex):
        pass

    def get_global_pose(self):
        return (self.entity.pose if self.entity is not None else Pose()) * self.local_pose

    global_pose = property(get_global_pose)

    def get_intrinsic_matrix(self):
        return np.array([[self._fx, self.skew, self._cx], [0, self._fy, self._cy], [0, 0, 1]], dtype=np.float32)

    def get_model_matrix(self):
        """cam2world in OpenGL camera axes (x right, y up, z back)."""
        T = self.get_global_pose().to_transformation_matrix()
        gl = np.array([[0, 0, -1, 0], [-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
        return (T @ gl).astype(np.float32)

    def get_extrinsic_matrix(self):
        """world2cam [3, 4] in OpenCV camera axes (x right, y down, z forward)."""
        T = self.get_global_pose().to_transformation_matrix()
        cv = np.array([[0, 0, 1, 0], [-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
        return np.linalg.inv(T @ cv)[:3].astype(np.float32)

    def get_projection_matrix(self):
        n, f = self.near, self.far
        P = np.zeros((4, 4), dtype=np.float32)
        P[0, 0], P[1, 1] = 2 * self._fx / self.width, 2 * self._fy / self.height
        P[0, 2], P[1, 2] = 1 - 2 * self._cx / self.width, 2 * self._cy / self.height - 1
        P[2, 2], P[2, 3], P[3, 2] = -(f + n) / (f - n), -2 * f * n / (f - n), -1
        return P

    # single-camera rendering is a CPU-simulation path
    def take_picture(self):
        raise NotImplementedError("single-camera rendering is a CPU-simulation call; cameras render through their camera group on this backend")

    get_picture = get_picture_cuda = take_picture

    def _on_add_to_scene(self, scene):
        if scene.render_system is not None:
            scene.render_system.cameras.append(self)


# ------------------------------------------------------------------------------------------------ systems
class RenderSystem:
    def __init__(self, device=None):
        from .. import Device
        self.device = device if isinstance(device, Device) or device is None else Device(str(device))
        self.scene = None
        self.render_bodies: List[RenderBodyComponent] = []
        self.cameras: List[RenderCameraComponent] = []
        self.lights: List[_RenderLight] = []
        self.ambient_light = [0.0, 0.0, 0.0]
        self.cubemap = None

    def get_render_bodies(self):
        return self.render_bodies

    def get_cameras(self):
        return self.cameras

    def get_lights(self):
        return self.lights

    def set_ambient_light(self, c):
        self.ambient_light = list(c)

    def get_ambient_light(self):
        return self.ambient_light

    def set_cubemap(self, c):
        self.cubemap = c

    def get_cubemap(self):
        return self.cubemap

    def step(self):
        pass


class _PictureCuda:
    def __init__(self, tensor_fn):
        self._fn = tensor_fn

    def torch(self):
        return self._fn()

    @property
    def shape(self):
        return tuple(self._fn().shape)


class RenderCameraGroup:
    """What `create_camera_group` returns: `take_picture()` renders every sub-scene's camera; `get_picture_cuda(name).torch()` is the
    `[N, H, W, 4]` render target."""

    def __init__(self, group, formats):
        self._group = group
        self._formats = dict(formats)

    def take_picture(self):
        self._group.take_picture()

    def get_picture_cuda(self, name: str):
        if name not in ("Color", "PositionSegmentation"):
            raise RuntimeError(f"texture '{name}' is not produced by the b200sim rasteriser (Color, PositionSegmentation)")

        def fetch():
            import torch
            t = self._group.get_picture_cuda(name, 0)
            fmt = self._formats.get(name, "")
            if fmt.endswith("sfloat"):   # "default" shader pack: Color float32 in [0, 1], PositionSegmentation float32 (metres, ids)
                if name == "Color":
                    return t.to(torch.float32) / 255.0
                out = t.to(torch.float32)
                out[..., :3] /= 1000.0
                return out
            return t

        return _PictureCuda(fetch)


class RenderSystemGroup:
    def __init__(self, systems: Sequence[RenderSystem]):
        self.systems = list(systems)
        self._poses = None

    def set_cuda_poses(self, cuda_array):
        self._poses = cuda_array

    def update_render(self):
        """Nothing to synchronise: the rasteriser reads body poses from `cuda_rigid_body_data` when a picture is taken."""

    def create_camera_group(self, cameras: Sequence[RenderCameraComponent], texture_names: Sequence[str]):
        from maniskill_b200.compat.compile import create_camera_group
        for n in texture_names:
            if n not in ("Color", "PositionSegmentation"):
                raise RuntimeError(f"texture '{n}' is not produced by the b200sim rasteriser (use the 'minimal' or 'default' shader pack)")
        return RenderCameraGroup(create_camera_group(self, list(cameras)), _PICTURE_FORMAT)


class RenderManager:
    def __init__(self, *a, **kw):
        raise NotImplementedError("sapien 3.1 render managers are not part of this surface (SAPIEN_RENDER_SYSTEM is '3.0')")


GpuSyncManager = RenderManager


def get_shader_pack(name):
    raise NotImplementedError("shader packs are a sapien 3.1 concept")
