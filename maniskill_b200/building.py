"""Scene-building surface of the reference on top of the scene tables: `sapien.Pose`, `ActorBuilder`, `actors.build_*`.

What a task's `_load_scene` calls in the reference (mani_skill/utils/building/actor_builder.py:20-300 -- a subclass of sapien's
`ActorBuilder`, whose record fields SURVEY.md section 8(b) lists: `collision_records` / `visual_records` with
`type, pose, scale, radius, length, material, density, patch_radius, min_patch_radius`, `physx_body_type`, `collision_groups`, `name`,
`initial_pose`; and the helpers of mani_skill/utils/building/actors/common.py).  Here the builder collects the same records and
`build()` turns them into one `ActorRec` of the `SceneDesc` prototype (maniskill_b200/model.py) that is compiled once and instantiated
`n_envs` times on the device -- instead of one PhysX entity per sub-scene.

    builder = scene_desc_builder(scene_desc)                  # scene.create_actor_builder()
    builder.add_box_collision(half_size=[0.02] * 3)
    builder.add_box_visual(half_size=[0.02] * 3, material=RenderMaterial(base_color=[1, 0, 0, 1]))
    builder.initial_pose = Pose(p=[0, 0, 0.02])
    cube = builder.build(name="cube")                          # -> the ActorRec added to the SceneDesc

Not supported (raises): mesh files (`add_*_from_file`; robot meshes are baked offline by tools/bake_assets.py), per-sub-scene builds
(`set_scene_idxs`; heterogeneous geometry goes through the per-env override tables of `ShapeRec`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .model import (SHAPE_BOX, SHAPE_CAPSULE, SHAPE_CONVEX, SHAPE_PLANE, SHAPE_SPHERE, ActorRec, ArticulationRec, SceneDesc, ShapeRec, combine_mass,
                    cylinder_shape, pose7, pose_inv, pose_mul, qmat, qrot, sym6)


class Device:
    """`sapien.Device("cuda" | "cuda:n" | "cpu")` (mani_skill/envs/utils/system/backend.py:60-91)."""

    def __init__(self, name: str):
        self.name = name
        kind, _, idx = name.partition(":")
        if kind not in ("cpu", "cuda"):
            raise ValueError(f"unknown device {name!r}")
        self.type, self.cuda_id = kind, (int(idx) if idx else 0)

    def is_cuda(self) -> bool:
        return self.type == "cuda"

    def is_cpu(self) -> bool:
        return self.type == "cpu"

    def __repr__(self):
        return f"Device({self.name!r})"


class Pose:
    """`sapien.Pose`: one rigid transform, numpy float32, quaternion wxyz (`.p .q`, `*`, `.inv()`,
    `.to_transformation_matrix()`, `Pose(4x4)`, `set_p/set_q`, `.rpy`)."""

    __slots__ = ("_v",)

    def __init__(self, p=(0, 0, 0), q=(1, 0, 0, 0)):
        p = np.asarray(p, dtype=np.float64)
        if p.shape == (4, 4):  # Pose(matrix)
            from scipy.spatial.transform import Rotation
            x, y, z, w = Rotation.from_matrix(p[:3, :3]).as_quat()
            self._v = pose7(p[:3, 3], [w, x, y, z])
            return
        q = np.asarray(q, dtype=np.float64)
        if p.shape != (3,) or q.shape != (4,):
            raise ValueError(f"Pose(p[3], q[4]) or Pose(matrix[4,4]); got shapes {p.shape}, {q.shape}")
        self._v = pose7(p, q)

    @classmethod
    def _from7(cls, v):
        out = cls.__new__(cls)
        out._v = np.asarray(v, dtype=np.float64).copy()
        return out

    @property
    def p(self):
        return self._v[:3].astype(np.float32)

    @property
    def q(self):
        return self._v[3:].astype(np.float32)

    def set_p(self, p):
        self._v[:3] = np.asarray(p, dtype=np.float64)

    def set_q(self, q):
        self._v[3:] = np.asarray(q, dtype=np.float64)

    def get_p(self):
        return self.p

    def get_q(self):
        return self.q

    @property
    def rpy(self):
        """Roll, pitch, yaw of the fixed-axis x-y-z convention (`transforms3d.euler.quat2euler` default 'sxyz')."""
        R = qmat(self._v[3:] / np.linalg.norm(self._v[3:]))
        return np.array([np.arctan2(R[2, 1], R[2, 2]), np.arcsin(-np.clip(R[2, 0], -1, 1)), np.arctan2(R[1, 0], R[0, 0])], dtype=np.float32)

    def __mul__(self, other: "Pose") -> "Pose":
        return Pose._from7(pose_mul(self._v, other._v))

    def inv(self) -> "Pose":
        return Pose._from7(pose_inv(self._v))

    def to_transformation_matrix(self):
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = qmat(self._v[3:] / np.linalg.norm(self._v[3:]))
        T[:3, 3] = self._v[:3]
        return T

    def raw(self):
        """[p, q] as the 7-vector the scene tables store."""
        return self._v.copy()

    def __repr__(self):
        return f"Pose({self.p.tolist()}, {self.q.tolist()})"


def _pose7_of(pose) -> np.ndarray:
    if pose is None:
        return pose7()
    if isinstance(pose, Pose):
        return pose.raw()
    v = np.asarray(pose, dtype=np.float64).reshape(-1)
    if v.shape != (7,):
        raise ValueError("pose must be a building.Pose or a 7-vector (p, q wxyz)")
    return v


@dataclass
class PhysxMaterial:
    """`sapien.physx.PhysxMaterial(static_friction, dynamic_friction, restitution)`; the solver uses one Coulomb coefficient per shape
    (the dynamic one; the reference's materials set both to the same value) and restitution 0."""
    static_friction: float = 0.3
    dynamic_friction: float = 0.3
    restitution: float = 0.0


@dataclass
class RenderMaterial:
    """`sapien.render.RenderMaterial`: the rasteriser shades flat base colours."""
    base_color: Sequence[float] = (0.7, 0.7, 0.7, 1.0)
    roughness: float = 0.5
    specular: float = 0.5
    metallic: float = 0.0

    def set_base_color(self, color: Sequence[float]):
        self.base_color = tuple(float(c) for c in color)


@dataclass
class CollisionRecord:
    type: str
    pose: np.ndarray
    scale: np.ndarray = field(default_factory=lambda: np.ones(3))   # box half sizes
    radius: float = 0.0
    length: float = 0.0                                             # half length of capsules / cylinders (axis = local x)
    material: Optional[PhysxMaterial] = None
    density: float = 1000.0
    patch_radius: float = 0.0
    min_patch_radius: float = 0.0


@dataclass
class VisualRecord:
    type: str
    pose: np.ndarray
    scale: np.ndarray = field(default_factory=lambda: np.ones(3))
    radius: float = 0.0
    length: float = 0.0
    material: Optional[RenderMaterial] = None
    name: str = ""


class ActorBuilder:
    """actor_builder.py:20-300.  `scene` is the SceneDesc the built actor is added to."""

    def __init__(self, scene: Optional[SceneDesc] = None):
        self.scene = scene
        self.collision_records: List[CollisionRecord] = []
        self.visual_records: List[VisualRecord] = []
        self.physx_body_type = "dynamic"
        self.collision_groups = [1, 1, 0, 0]
        self.name = ""
        self.initial_pose = None
        self._auto_inertial = True
        self._mass = self._cmass_local_pose = self._inertia = None
        self._hidden = False

    # ---- setters of the sapien base class
    def set_scene(self, scene: SceneDesc):
        self.scene = scene
        return self

    def set_name(self, name: str):
        self.name = name
        return self

    def set_physx_body_type(self, t: str):
        if t not in ("dynamic", "kinematic", "static"):
            raise Exception(f"invalid physx body type [{t}]")   # actor_builder.py:78 words it the same way
        self.physx_body_type = t
        return self

    def set_initial_pose(self, pose):
        self.initial_pose = pose
        return self

    def set_collision_groups(self, groups: Sequence[int]):
        if len(groups) != 4:
            raise ValueError("collision groups are 4 x uint32 (actor_builder.py:151)")
        self.collision_groups = [int(g) for g in groups]
        return self

    def set_mass_and_inertia(self, mass: float, cmass_local_pose, inertia: Sequence[float]):
        """Explicit inertial: mass, centre-of-mass frame (its rotation = the principal axes) and the principal moments."""
        self._auto_inertial = False
        self._mass, self._cmass_local_pose, self._inertia = float(mass), _pose7_of(cmass_local_pose), np.asarray(inertia, dtype=np.float64)
        return self

    def set_scene_idxs(self, scene_idxs=None):
        if scene_idxs is not None:
            raise NotImplementedError("one prototype is instantiated in every sub-scene; per-env geometry goes through ShapeRec.per_env_size / per_env_pose")
        return self

    # ---- collision records
    def _add_collision(self, **kw):
        self.collision_records.append(CollisionRecord(**kw))
        return self

    def add_box_collision(self, pose=None, half_size=(1, 1, 1), material=None, density=1000.0, patch_radius=0.0, min_patch_radius=0.0):
        return self._add_collision(type="box", pose=_pose7_of(pose), scale=np.asarray(half_size, dtype=np.float64), material=material, density=density,
                                   patch_radius=patch_radius, min_patch_radius=min_patch_radius)

    def add_sphere_collision(self, pose=None, radius=1.0, material=None, density=1000.0, patch_radius=0.0, min_patch_radius=0.0):
        return self._add_collision(type="sphere", pose=_pose7_of(pose), radius=float(radius), material=material, density=density,
                                   patch_radius=patch_radius, min_patch_radius=min_patch_radius)

    def add_capsule_collision(self, pose=None, radius=1.0, half_length=1.0, material=None, density=1000.0, patch_radius=0.0, min_patch_radius=0.0):
        return self._add_collision(type="capsule", pose=_pose7_of(pose), radius=float(radius), length=float(half_length), material=material, density=density,
                                   patch_radius=patch_radius, min_patch_radius=min_patch_radius)

    def add_cylinder_collision(self, pose=None, radius=1.0, half_length=1.0, material=None, density=1000.0, patch_radius=0.0, min_patch_radius=0.0):
        return self._add_collision(type="cylinder", pose=_pose7_of(pose), radius=float(radius), length=float(half_length), material=material, density=density,
                                   patch_radius=patch_radius, min_patch_radius=min_patch_radius)

    def add_plane_collision(self, pose=None, material=None):
        """Infinite plane, normal = local +x (sapien convention; scene_builder/table uses q = [0.7071, 0, -0.7071, 0] for a floor)."""
        return self._add_collision(type="plane", pose=_pose7_of(pose), material=material)

    def add_convex_collision_from_file(self, *a, **k):
        raise NotImplementedError("mesh files are cooked offline (tools/bake_assets.py); build the ShapeRec from its vertices instead")

    add_multiple_convex_collisions_from_file = add_nonconvex_collision_from_file = add_visual_from_file = add_convex_collision_from_file

    # ---- visual records
    def _add_visual(self, **kw):
        self.visual_records.append(VisualRecord(**kw))
        return self

    def add_box_visual(self, pose=None, half_size=(1, 1, 1), material=None, name=""):
        return self._add_visual(type="box", pose=_pose7_of(pose), scale=np.asarray(half_size, dtype=np.float64), material=material, name=name)

    def add_sphere_visual(self, pose=None, radius=1.0, material=None, name=""):
        return self._add_visual(type="sphere", pose=_pose7_of(pose), radius=float(radius), material=material, name=name)

    def add_capsule_visual(self, pose=None, radius=1.0, half_length=1.0, material=None, name=""):
        return self._add_visual(type="capsule", pose=_pose7_of(pose), radius=float(radius), length=float(half_length), material=material, name=name)

    def add_cylinder_visual(self, pose=None, radius=1.0, half_length=1.0, material=None, name=""):
        return self._add_visual(type="cylinder", pose=_pose7_of(pose), radius=float(radius), length=float(half_length), material=material, name=name)

    # ---- build
    @staticmethod
    def _geometry(r):
        """(shape type, size, vertices, triangles) of a record."""
        if r.type == "box":
            return SHAPE_BOX, np.asarray(r.scale, dtype=np.float64), None, None
        if r.type == "sphere":
            return SHAPE_SPHERE, np.array([r.radius, 0.0, 0.0]), None, None
        if r.type == "capsule":
            return SHAPE_CAPSULE, np.array([r.radius, r.length, 0.0]), None, None
        if r.type == "cylinder":
            c = cylinder_shape(r.radius, r.length)
            return SHAPE_CONVEX, np.zeros(3), c.vertices, c.triangles
        if r.type == "plane":
            return SHAPE_PLANE, np.zeros(3), None, None
        raise RuntimeError(f"invalid collision shape type [{r.type}]")   # actor_builder.py:138

    def _shape_recs(self) -> List[ShapeRec]:
        shapes: List[ShapeRec] = []
        visuals = list(self.visual_records)
        for r in self.collision_records:
            t, size, verts, tris = self._geometry(r)
            mat = r.material or PhysxMaterial()
            # a visual record of the same geometry at the same place is the same ShapeRec (one entry in both the shape and the visual table)
            twin = next((v for v in visuals if v.type == r.type and np.allclose(v.pose, r.pose) and np.allclose(v.scale, r.scale)
                         and v.radius == r.radius and v.length == r.length), None)
            kw = dict(color=tuple((twin.material or RenderMaterial()).base_color)) if twin is not None else dict(visual=False)
            if twin is not None:
                visuals.remove(twin)
            shapes.append(ShapeRec(t, r.pose.copy(), size, vertices=verts, triangles=tris, mu=float(mat.dynamic_friction), patch_radius=float(r.patch_radius),
                                   density=float(r.density), groups=tuple(self.collision_groups), **kw))
        for v in visuals:
            t, size, verts, tris = self._geometry(v)
            shapes.append(ShapeRec(t, v.pose.copy(), size, vertices=verts, triangles=tris, collide=False,
                                   color=tuple((v.material or RenderMaterial()).base_color)))
        return shapes

    def build(self, name: Optional[str] = None) -> ActorRec:
        if name is not None:
            self.name = name
        if not self.name:
            raise ValueError("actor needs a name")   # the reference asserts non-empty unique names (actor_builder.py:216-222)
        if self.scene is None:
            raise RuntimeError("builder has no scene: use scene_desc_builder(scene_desc) or set_scene()")
        if any(a.name == self.name for a in self.scene.actors):
            raise RuntimeError(f"actor name {self.name!r} already used in this scene")
        rec = ActorRec(self.name, self.physx_body_type, self._shape_recs(), _pose7_of(self.initial_pose), hidden=self._hidden)
        if not self._auto_inertial and self.physx_body_type != "kinematic":   # actor_builder.py:156-160
            R = qmat(self._cmass_local_pose[3:])
            rec.mass, rec.com, rec.inertia = self._mass, self._cmass_local_pose[:3].copy(), R @ np.diag(self._inertia) @ R.T
        self.scene.add_actor(rec)
        return rec

    def build_kinematic(self, name: Optional[str] = None) -> ActorRec:
        self.physx_body_type = "kinematic"
        return self.build(name)

    def build_static(self, name: Optional[str] = None) -> ActorRec:
        self.physx_body_type = "static"
        return self.build(name)

    def build_dynamic(self, name: Optional[str] = None) -> ActorRec:
        self.physx_body_type = "dynamic"
        return self.build(name)


def scene_desc_builder(scene: SceneDesc) -> ActorBuilder:
    """`scene.create_actor_builder()` (mani_skill/envs/scene.py:183-190)."""
    return ActorBuilder(scene)


# ------------------------------------------------------------------------------------------------ mani_skill/utils/building/actors/common.py
def _build_by_type(builder: ActorBuilder, name, body_type, initial_pose=None):
    """common.py:22-49."""
    if initial_pose is not None:
        builder.set_initial_pose(initial_pose)
    if body_type not in ("dynamic", "static", "kinematic"):
        raise ValueError(f"Unknown body type {body_type}")
    return builder.set_physx_body_type(body_type).build(name=name)


def build_cube(scene: SceneDesc, half_size: float, color, name: str, body_type="dynamic", add_collision=True, initial_pose=None):
    """common.py:52-79."""
    b = scene_desc_builder(scene)
    if add_collision:
        b.add_box_collision(half_size=[half_size] * 3)
    b.add_box_visual(half_size=[half_size] * 3, material=RenderMaterial(base_color=color))
    return _build_by_type(b, name, body_type, initial_pose)


def build_box(scene: SceneDesc, half_sizes, color, name: str, body_type="dynamic", add_collision=True, initial_pose=None):
    """common.py:82-109."""
    b = scene_desc_builder(scene)
    if add_collision:
        b.add_box_collision(half_size=half_sizes)
    b.add_box_visual(half_size=half_sizes, material=RenderMaterial(base_color=color))
    return _build_by_type(b, name, body_type, initial_pose)


def build_sphere(scene: SceneDesc, radius: float, color, name: str, body_type="dynamic", add_collision=True, initial_pose=None):
    """common.py:145-166."""
    b = scene_desc_builder(scene)
    if add_collision:
        b.add_sphere_collision(radius=radius)
    b.add_sphere_visual(radius=radius, material=RenderMaterial(base_color=color))
    return _build_by_type(b, name, body_type, initial_pose)


def build_cylinder(scene: SceneDesc, radius: float, half_length: float, color, name: str, body_type="dynamic", add_collision=True, initial_pose=None):
    """common.py:112-142."""
    b = scene_desc_builder(scene)
    if add_collision:
        b.add_cylinder_collision(radius=radius, half_length=half_length)
    b.add_cylinder_visual(radius=radius, half_length=half_length, material=RenderMaterial(base_color=color))
    return _build_by_type(b, name, body_type, initial_pose)


def build_twocolor_peg(scene: SceneDesc, length, width, color_1, color_2, name: str, body_type="dynamic", add_collision=True, initial_pose=None):
    """common.py:230-261 (`length`, `width` are half extents)."""
    b = scene_desc_builder(scene)
    if add_collision:
        b.add_box_collision(half_size=[length, width, width])
    b.add_box_visual(pose=Pose(p=[-length / 2, 0, 0]), half_size=[length / 2, width, width], material=RenderMaterial(base_color=color_1))
    b.add_box_visual(pose=Pose(p=[length / 2, 0, 0]), half_size=[length / 2, width, width], material=RenderMaterial(base_color=color_2))
    return _build_by_type(b, name, body_type, initial_pose)


# ------------------------------------------------------------------------------------------------ articulations
@dataclass
class JointRecord:
    """articulation_builder.py / sapien `LinkBuilder.joint_record`: the joint frame is given on both sides, the joint
    moves about / along the frame's x axis."""
    joint_type: str = "undefined"
    limits: Sequence[float] = (-np.inf, np.inf)
    pose_in_parent: np.ndarray = field(default_factory=pose7)
    pose_in_child: np.ndarray = field(default_factory=pose7)
    friction: float = 0.0
    damping: float = 0.0
    name: str = ""


class LinkBuilder(ActorBuilder):
    """`sapien.wrapper.articulation_builder.LinkBuilder(index, parent)`: an actor builder (collision / visual records of the link) + the
    record of the joint that connects it to its parent."""

    def __init__(self, index: int, parent: Optional["LinkBuilder"] = None):
        super().__init__(None)
        self.index, self.parent = index, parent
        self.physx_body_type = "link"
        self.joint_record = JointRecord()

    def set_joint_name(self, name: str):
        self.joint_record.name = name
        return self

    def set_joint_properties(self, type: str, limits, pose_in_parent=None, pose_in_child=None, friction: float = 0.0, damping: float = 0.0):
        if type not in ("fixed", "revolute", "revolute_unwrapped", "continuous", "prismatic", "undefined"):
            raise ValueError(f"unknown joint type {type!r}")
        lim = np.asarray(limits, dtype=np.float64).reshape(-1) if len(np.asarray(limits).reshape(-1)) else np.array([-np.inf, np.inf])
        j = self.joint_record
        j.joint_type, j.limits, j.friction, j.damping = type, lim, float(friction), float(damping)
        j.pose_in_parent, j.pose_in_child = _pose7_of(pose_in_parent), _pose7_of(pose_in_child)
        return self

    def _check(self):
        """articulation_builder's `_check`: every non-root link needs a joint."""
        if self.parent is not None and self.joint_record.joint_type == "undefined":
            raise RuntimeError(f"link {self.name!r} has a parent but no joint properties")

    def build(self, name=None):
        raise RuntimeError("links are built by their ArticulationBuilder.build()")


class ArticulationBuilder:
    """mani_skill/utils/building/articulation_builder.py:23-215 on the scene tables: `create_link_builder(parent)` per link, then
    `build(name, fix_root_link=True)` compiles the records into one robot description (the format tools/bake_assets.py bakes URDFs
    into) and adds an `ArticulationRec` to the SceneDesc.  Fixed-base articulations only (the device integrates no floating roots)."""

    def __init__(self, scene: Optional[SceneDesc] = None):
        self.scene = scene
        self.link_builders: List[LinkBuilder] = []
        self.mimic_joint_records: List[dict] = []
        self.name = ""
        self.initial_pose = None
        self.disable_gravity = False

    def set_scene(self, scene: SceneDesc):
        self.scene = scene
        return self

    def set_name(self, name: str):
        self.name = name
        return self

    def set_initial_pose(self, pose):
        self.initial_pose = pose
        return self

    def create_link_builder(self, parent: Optional[LinkBuilder] = None) -> LinkBuilder:
        if parent is not None and parent not in self.link_builders:
            raise ValueError("parent link builder belongs to another articulation")
        if parent is None and self.link_builders:
            raise ValueError("an articulation has one root link")
        b = LinkBuilder(len(self.link_builders), parent)
        self.link_builders.append(b)
        return b

    def add_mimic_joint(self, joint: str, mimic: str, multiplier: float = 1.0, offset: float = 0.0):
        """`mimic_joint_records`: joint = multiplier * mimic + offset (a fixed tendon in the reference, articulation_builder.py:161-200)."""
        self.mimic_joint_records.append(dict(joint=joint, mimic=mimic, multiplier=float(multiplier), offset=float(offset)))
        return self

    def _robot(self) -> dict:
        links = []
        for b in self.link_builders:
            b._check()
            if not b.name:
                raise ValueError("every link needs a name")
            shapes = b._shape_recs()
            if any(s.type in (SHAPE_CAPSULE, SHAPE_PLANE) and s.collide for s in shapes):
                raise NotImplementedError("link collision shapes: box, sphere, cylinder (articulated capsules / planes are not in the link shape table)")
            cols = []
            for s in shapes:
                if not s.collide:
                    continue   # link visuals follow the collision geometry in this renderer
                c = dict(p=s.pose[:3].tolist(), q=s.pose[3:].tolist())
                if s.type == SHAPE_BOX:
                    c.update(type="box", half_size=np.asarray(s.size).tolist())
                elif s.type == SHAPE_SPHERE:
                    c.update(type="sphere", radius=float(s.size[0]))
                else:
                    c.update(type="convex", vertices=np.asarray(s.vertices).tolist(), triangles=np.asarray(s.triangles).tolist())
                cols.append(c)
            if b._auto_inertial:
                m, com, I = combine_mass([s.mass_props() for s in shapes if s.collide])
            else:
                R = qmat(b._cmass_local_pose[3:])
                m, com, I = b._mass, b._cmass_local_pose[:3], R @ np.diag(b._inertia) @ R.T
            j = b.joint_record
            link = dict(name=b.name, parent=-1 if b.parent is None else b.parent.index, mass=float(m), com=np.asarray(com).tolist(),
                        inertia=[float(x) for x in sym6(np.asarray(I))], collisions=cols)
            if b.parent is None or j.joint_type in ("fixed", "undefined"):
                T0 = pose_mul(j.pose_in_parent, pose_inv(j.pose_in_child)) if b.parent is not None else pose7()
                link["joint"] = dict(name=j.name, type="fixed", p=T0[:3].tolist(), q=T0[3:].tolist(), axis=[1, 0, 0], lower=0, upper=0, effort=0,
                                     damping=0, friction=0)
            else:
                C = j.pose_in_child
                rot_c_inv = pose7([0, 0, 0], [C[3], -C[4], -C[5], -C[6]])
                T0 = pose_mul(j.pose_in_parent, rot_c_inv)                      # frame the joint moves: anchored at the joint, link orientation
                axis = qrot(C[3:], [1.0, 0.0, 0.0])
                jt = "revolute" if j.joint_type == "continuous" else j.joint_type
                lo, hi = (float(j.limits[0]), float(j.limits[1])) if j.joint_type != "continuous" else (-1e30, 1e30)
                lo, hi = max(lo, -1e30), min(hi, 1e30)
                link["joint"] = dict(name=j.name, type=jt, p=T0[:3].tolist(), q=T0[3:].tolist(), axis=axis.tolist(), lower=lo, upper=hi, effort=0,
                                     damping=j.damping, friction=j.friction)
                if np.abs(C[:3]).max() > 0:
                    link["frame_offset"] = pose7(-C[:3]).tolist()               # link frame inside that moving frame
            links.append(link)
        by_joint = {l["joint"]["name"]: l for l in links if l["joint"]["name"]}
        for r in self.mimic_joint_records:
            if r["joint"] not in by_joint or r["mimic"] not in by_joint:
                raise ValueError(f"mimic record names unknown joints: {r}")
            by_joint[r["joint"]]["joint"]["mimic"] = dict(joint=r["mimic"], multiplier=r["multiplier"], offset=r["offset"])
        return dict(name=self.name, source="ArticulationBuilder", links=links, disabled_collision_pairs=[])

    def build(self, name: Optional[str] = None, fix_root_link: Optional[bool] = True, build_mimic_joints: bool = True) -> ArticulationRec:
        if name is not None:
            self.name = name
        if self.scene is None:
            raise RuntimeError("builder has no scene")
        if not self.name or any(a.name == self.name for a in self.scene.articulations):
            raise RuntimeError("built articulations must have unique, non-empty names")   # articulation_builder.py:120-126
        if not fix_root_link:
            raise NotImplementedError("floating-base articulations are not supported: build with fix_root_link=True")
        if not self.link_builders:
            raise ValueError("articulation without links")
        if not build_mimic_joints:
            self.mimic_joint_records = []
        rec = ArticulationRec(self.name, self._robot(), _pose7_of(self.initial_pose), disable_gravity=self.disable_gravity)
        rec.joint_friction = {b.joint_record.name: b.joint_record.friction for b in self.link_builders if b.joint_record.name}
        rec.link_groups = {b.name: tuple(b.collision_groups) for b in self.link_builders}
        rec.link_mu = {b.name: float((b.collision_records[0].material or PhysxMaterial()).dynamic_friction) for b in self.link_builders if b.collision_records}
        rec.link_patch = {b.name: float(b.collision_records[0].patch_radius) for b in self.link_builders if b.collision_records}
        self.scene.add_articulation(rec)
        return rec


def scene_desc_articulation_builder(scene: SceneDesc) -> ArticulationBuilder:
    """`scene.create_articulation_builder()` (mani_skill/envs/scene.py:192-199)."""
    return ArticulationBuilder(scene)
