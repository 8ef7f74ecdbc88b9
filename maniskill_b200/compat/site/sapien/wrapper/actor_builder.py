"""sapien.wrapper.actor_builder: the `ActorBuilder` base class the reference subclasses (mani_skill/utils/building/actor_builder.py:20) --
collision / visual records with the field names SURVEY.md 8(b) lists, and the component factories `build_physx_component` /
`build_render_component` that turn the records into (recording) components."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

import sapien
from sapien import physx, render


def _mat(material):
    return material if material is not None else physx.get_default_material()


def _rmat(material):
    if material is None:
        return render.RenderMaterial()
    if isinstance(material, render.RenderMaterial):
        return material
    c = list(material)
    return render.RenderMaterial(base_color=c + [1.0] * (4 - len(c)))


@dataclass
class CollisionShapeRecord:
    type: str
    filename: str = ""
    scale: tuple = (1, 1, 1)          # box: half sizes
    radius: float = 1
    length: float = 1                 # half length of capsules / cylinders (axis = local x)
    material: Optional[physx.PhysxMaterial] = None
    pose: sapien.Pose = field(default_factory=sapien.Pose)
    density: float = 1000
    patch_radius: float = 0
    min_patch_radius: float = 0
    is_trigger: bool = False
    decomposition: str = "none"
    decomposition_params: Optional[dict] = None


@dataclass
class VisualShapeRecord:
    type: str
    filename: str = ""
    scale: tuple = (1, 1, 1)
    radius: float = 1
    length: float = 1
    material: Optional[render.RenderMaterial] = None
    pose: sapien.Pose = field(default_factory=sapien.Pose)
    name: str = ""


class ActorBuilder:
    def __init__(self):
        self.collision_records: List[CollisionShapeRecord] = []
        self.visual_records: List[VisualShapeRecord] = []
        self.use_density = True
        self.collision_groups = [1, 1, 0, 0]
        self.scene = None
        self.physx_body_type = "dynamic"
        self.name = ""
        self.initial_pose = sapien.Pose()
        self._mass = 1.0
        self._inertia = np.zeros(3)
        self._cmass_local_pose = sapien.Pose()
        self._auto_inertial = True

    # ---- fluent setters
    def set_scene(self, scene):
        self.scene = scene
        return self

    def set_name(self, name):
        self.name = name
        return self

    def set_initial_pose(self, pose):
        self.initial_pose = pose
        return self

    def set_physx_body_type(self, type):
        if type not in ("dynamic", "kinematic", "static", "link"):
            raise Exception(f"invalid physx body type [{type}]")
        self.physx_body_type = type
        return self

    def set_collision_groups(self, group0, group1=None, group2=None, group3=None):
        g = list(group0) if group1 is None else [group0, group1, group2, group3]
        self.collision_groups = [int(x) for x in g]
        return self

    def set_mass_and_inertia(self, mass, cmass_local_pose, inertia):
        self._mass, self._cmass_local_pose, self._inertia = float(mass), cmass_local_pose, np.asarray(inertia, dtype=np.float64)
        self._auto_inertial = False
        return self

    def reset_mass_and_inertia(self):
        self._auto_inertial = True
        return self

    # ---- collision records
    def _col(self, **kw):
        self.collision_records.append(CollisionShapeRecord(**kw))
        return self

    def add_plane_collision(self, pose=None, material=None, patch_radius=0, min_patch_radius=0, is_trigger=False):
        return self._col(type="plane", pose=pose or sapien.Pose(), material=_mat(material), patch_radius=patch_radius, min_patch_radius=min_patch_radius, is_trigger=is_trigger)

    def add_box_collision(self, pose=None, half_size=(1, 1, 1), material=None, density=1000, patch_radius=0, min_patch_radius=0, is_trigger=False):
        return self._col(type="box", pose=pose or sapien.Pose(), scale=tuple(float(x) for x in half_size), material=_mat(material), density=density,
                         patch_radius=patch_radius, min_patch_radius=min_patch_radius, is_trigger=is_trigger)

    def add_capsule_collision(self, pose=None, radius=1, half_length=1, material=None, density=1000, patch_radius=0, min_patch_radius=0, is_trigger=False):
        return self._col(type="capsule", pose=pose or sapien.Pose(), radius=radius, length=half_length, material=_mat(material), density=density,
                         patch_radius=patch_radius, min_patch_radius=min_patch_radius, is_trigger=is_trigger)

    def add_cylinder_collision(self, pose=None, radius=1, half_length=1, material=None, density=1000, patch_radius=0, min_patch_radius=0, is_trigger=False):
        return self._col(type="cylinder", pose=pose or sapien.Pose(), radius=radius, length=half_length, material=_mat(material), density=density,
                         patch_radius=patch_radius, min_patch_radius=min_patch_radius, is_trigger=is_trigger)

    def add_sphere_collision(self, pose=None, radius=1, material=None, density=1000, patch_radius=0, min_patch_radius=0, is_trigger=False):
        return self._col(type="sphere", pose=pose or sapien.Pose(), radius=radius, material=_mat(material), density=density, patch_radius=patch_radius,
                         min_patch_radius=min_patch_radius, is_trigger=is_trigger)

    def add_convex_collision_from_file(self, filename, pose=None, scale=(1, 1, 1), material=None, density=1000, patch_radius=0, min_patch_radius=0, is_trigger=False):
        return self._col(type="convex_mesh", filename=str(filename), pose=pose or sapien.Pose(), scale=tuple(np.broadcast_to(scale, (3,)).tolist()), material=_mat(material),
                         density=density, patch_radius=patch_radius, min_patch_radius=min_patch_radius, is_trigger=is_trigger)

    def add_multiple_convex_collisions_from_file(self, filename, pose=None, scale=(1, 1, 1), material=None, density=1000, patch_radius=0, min_patch_radius=0,
                                                 is_trigger=False, decomposition="none", decomposition_params=None):
        return self._col(type="multiple_convex_meshes", filename=str(filename), pose=pose or sapien.Pose(), scale=tuple(np.broadcast_to(scale, (3,)).tolist()),
                         material=_mat(material), density=density, patch_radius=patch_radius, min_patch_radius=min_patch_radius, is_trigger=is_trigger,
                         decomposition=decomposition, decomposition_params=decomposition_params or dict())

    def add_nonconvex_collision_from_file(self, filename, pose=None, scale=(1, 1, 1), material=None, patch_radius=0, min_patch_radius=0, is_trigger=False):
        return self._col(type="nonconvex_mesh", filename=str(filename), pose=pose or sapien.Pose(), scale=tuple(np.broadcast_to(scale, (3,)).tolist()), material=_mat(material),
                         patch_radius=patch_radius, min_patch_radius=min_patch_radius, is_trigger=is_trigger)

    # ---- visual records
    def _vis(self, **kw):
        self.visual_records.append(VisualShapeRecord(**kw))
        return self

    def add_plane_visual(self, pose=None, scale=(1, 1, 1), material=None, name=""):
        return self._vis(type="plane", pose=pose or sapien.Pose(), scale=tuple(scale), material=_rmat(material), name=name)

    def add_box_visual(self, pose=None, half_size=(1, 1, 1), material=None, name=""):
        return self._vis(type="box", pose=pose or sapien.Pose(), scale=tuple(float(x) for x in half_size), material=_rmat(material), name=name)

    def add_capsule_visual(self, pose=None, radius=1, half_length=1, material=None, name=""):
        return self._vis(type="capsule", pose=pose or sapien.Pose(), radius=radius, length=half_length, material=_rmat(material), name=name)

    def add_cylinder_visual(self, pose=None, radius=1, half_length=1, material=None, name=""):
        return self._vis(type="cylinder", pose=pose or sapien.Pose(), radius=radius, length=half_length, material=_rmat(material), name=name)

    def add_sphere_visual(self, pose=None, radius=1, material=None, name=""):
        return self._vis(type="sphere", pose=pose or sapien.Pose(), radius=radius, material=_rmat(material), name=name)

    def add_visual_from_file(self, filename, pose=None, scale=(1, 1, 1), material=None, name=""):
        return self._vis(type="file", filename=str(filename), pose=pose or sapien.Pose(), scale=tuple(np.broadcast_to(scale, (3,)).tolist()),
                         material=None if material is None else _rmat(material), name=name)

    # ---- component factories
    def build_render_component(self):
        component = render.RenderBodyComponent()
        for r in self.visual_records:
            if r.type == "plane":
                shape = render.RenderShapePlane(r.scale, r.material)
            elif r.type == "box":
                shape = render.RenderShapeBox(r.scale, r.material)
            elif r.type == "sphere":
                shape = render.RenderShapeSphere(r.radius, r.material)
            elif r.type == "capsule":
                shape = render.RenderShapeCapsule(r.radius, r.length, r.material)
            elif r.type == "cylinder":
                shape = render.RenderShapeCylinder(r.radius, r.length, r.material)
            elif r.type == "file":
                shape = render.RenderShapeTriangleMesh(r.filename, r.scale, r.material)
            else:
                raise Exception(f"invalid visual shape type [{r.type}]")
            shape.local_pose = r.pose
            shape.name = r.name
            component.attach(shape)
        return component

    def build_physx_component(self, link_parent=None):
        kind = self.physx_body_type
        if kind == "dynamic":
            component = physx.PhysxRigidDynamicComponent()
        elif kind == "kinematic":
            component = physx.PhysxRigidDynamicComponent()
            component.kinematic = True
        elif kind == "static":
            component = physx.PhysxRigidStaticComponent()
        elif kind == "link":
            component = physx.PhysxArticulationLinkComponent(link_parent)
        else:
            raise Exception(f"invalid physx body type [{kind}]")
        for r in self.collision_records:
            try:
                shapes = _collision_shapes(r)
            except RuntimeError:
                continue  # e.g. a mesh that cannot be cooked
            for shape in shapes:
                shape.local_pose = r.pose
                shape.set_collision_groups(self.collision_groups)
                shape.set_density(r.density)
                shape.set_patch_radius(r.patch_radius)
                shape.set_min_patch_radius(r.min_patch_radius)
                component.attach(shape)
        if not self._auto_inertial and kind != "kinematic":
            component.mass = self._mass
            component.cmass_local_pose = self._cmass_local_pose
            component.inertia = self._inertia
        if hasattr(self, "_srdf_disabled") and kind == "link":
            component.articulation.srdf_disabled_pairs = list(self._srdf_disabled)
        component.name = self.name
        return component

    def build_entity(self):
        entity = sapien.Entity()
        if self.visual_records:
            entity.add_component(self.build_render_component())
        entity.add_component(self.build_physx_component())
        entity.name = self.name
        return entity

    def build(self, name=None):
        if name is not None:
            self.set_name(name)
        entity = self.build_entity()
        entity.pose = self.initial_pose if self.initial_pose is not None else sapien.Pose()
        if self.scene is not None:
            self.scene.add_entity(entity)
        return entity

    def build_kinematic(self, name=""):
        return self.set_physx_body_type("kinematic").build(name=name)

    def build_static(self, name=""):
        return self.set_physx_body_type("static").build(name=name)


def _collision_shapes(r: CollisionShapeRecord):
    if r.type == "plane":
        return [physx.PhysxCollisionShapePlane(material=r.material)]
    if r.type == "box":
        return [physx.PhysxCollisionShapeBox(half_size=r.scale, material=r.material)]
    if r.type == "capsule":
        return [physx.PhysxCollisionShapeCapsule(radius=r.radius, half_length=r.length, material=r.material)]
    if r.type == "cylinder":
        return [physx.PhysxCollisionShapeCylinder(radius=r.radius, half_length=r.length, material=r.material)]
    if r.type == "sphere":
        return [physx.PhysxCollisionShapeSphere(radius=r.radius, material=r.material)]
    if r.type == "convex_mesh":
        return [physx.PhysxCollisionShapeConvexMesh(filename=r.filename, scale=r.scale, material=r.material)]
    if r.type == "nonconvex_mesh":
        return [physx.PhysxCollisionShapeTriangleMesh(filename=r.filename, scale=r.scale, material=r.material)]
    if r.type == "multiple_convex_meshes":
        from .coacd import do_coacd
        filename = do_coacd(r.filename, **(r.decomposition_params or {})) if r.decomposition == "coacd" else r.filename
        return physx.PhysxCollisionShapeConvexMesh.load_multiple(filename=filename, scale=r.scale, material=r.material)
    raise RuntimeError(f"invalid collision shape type [{r.type}]")
