"""`sapien` for the unmodified reference (haosulab/ManiSkill) on the b200sim backend -- SURVEY.md section 8(b) B1.

ManiSkill is pure python over the external `sapien` package; this package provides the surface it touches (exhaustive list in
SURVEY 8(b)) so that `import mani_skill; gym.make("PickCube-v1", num_envs=N)` runs byte-for-byte reference code:

* scene graph objects (`Scene`, `Entity`, components, collision / render shapes, articulations, joints) are lightweight RECORDING
  objects: building a scene only collects what the reference's builders set;
* `physx.PhysxGpuSystem.gpu_init()` compiles sub-scene 0 as the prototype (+ per-sub-scene shape sizes / poses where the
  sub-scenes differ) into the model tables of maniskill_b200/model.py and creates ONE batched world through the C-ABI
  (include/b200sim.h); from then on every `px.*` call is one C-ABI call (maniskill_b200/physx_shim.py);
* `render.RenderSystemGroup.create_camera_group` maps onto the batched rasteriser (b2s_camera_group_create).

Installed by maniskill_b200.compat.install() only when no real `sapien` is importable.  There is no CPU simulation path:
`PhysxCpuSystem` raises (tests inject an emulated world through maniskill_b200.compat.WORLD_FACTORY).
"""
from __future__ import annotations

import sys as _sys

import numpy as _np

from maniskill_b200.building import Pose as _BPose
from maniskill_b200.building import Device  # noqa: F401

__version__ = "3.0.0.b200sim"


class Pose(_BPose):
    """`sapien.Pose(p=[3], q=[4 wxyz])`: numpy float32 `.p` / `.q` (settable), `*`, `.inv()`, `.to_transformation_matrix()`, `Pose(4x4)`."""
    __slots__ = ()

    def __init__(self, p=(0, 0, 0), q=(1, 0, 0, 0)):
        p = _np.asarray(p, dtype=_np.float64)
        if p.shape != (4, 4):
            p = p.reshape(-1)
            q = _np.asarray(q, dtype=_np.float64).reshape(-1)
        super().__init__(p, q)

    p = property(_BPose.p.fget, lambda self, v: self.set_p(_np.asarray(v, dtype=_np.float64).reshape(3)))
    q = property(_BPose.q.fget, lambda self, v: self.set_q(_np.asarray(v, dtype=_np.float64).reshape(4)))

    def __mul__(self, other):
        out = _BPose.__mul__(self, other)
        return Pose._from7(out.raw())

    def inv(self):
        return Pose._from7(_BPose.inv(self).raw())

    def get_rpy(self):
        return self.rpy

    def set_rpy(self, rpy):
        from transforms3d.euler import euler2quat
        self.set_q(euler2quat(*rpy))

    def __getstate__(self):
        return self.raw()

    def __setstate__(self, state):
        self._v = _np.asarray(state, dtype=_np.float64).copy()

    def __copy__(self):
        return Pose._from7(self.raw())

    def __deepcopy__(self, memo):
        return Pose._from7(self.raw())


class Component:
    """Base of everything attached to an `Entity`."""
    name = ""

    def __init__(self):
        self.entity = None
        self.name = ""

    def get_entity(self):
        return self.entity

    def get_name(self):
        return self.name

    def set_name(self, name):
        self.name = name

    # pose of the owning entity (several component kinds expose it)
    @property
    def entity_pose(self):
        return self.entity.pose if self.entity is not None else Pose()

    def get_entity_pose(self):
        return self.entity_pose

    def _on_add_to_scene(self, scene):
        pass

    def _on_remove_from_scene(self, scene):
        pass


class Entity:
    """`sapien.Entity`: a named pose with components (actor_builder.py:178-191)."""

    def __init__(self):
        self.name = ""
        self._pose = Pose()
        self.components = []
        self.per_scene_id = 0
        self.scene = None

    # -- pose (also drives the physics body of the entity, like sapien does)
    @property
    def pose(self):
        for c in self.components:
            p = c._body_pose() if hasattr(c, "_body_pose") else None
            if p is not None:
                return p
        return self._pose

    @pose.setter
    def pose(self, pose):
        self._pose = Pose._from7(pose.raw())
        for c in self.components:
            if hasattr(c, "_set_body_pose"):
                c._set_body_pose(self._pose)

    def get_pose(self):
        return self.pose

    def set_pose(self, pose):
        self.pose = pose

    def get_name(self):
        return self.name

    def set_name(self, name):
        self.name = name

    def get_components(self):
        return self.components

    def get_scene(self):
        return self.scene

    def get_per_scene_id(self):
        return self.per_scene_id

    def add_component(self, component):
        component.entity = self
        self.components.append(component)
        if self.scene is not None:
            component._on_add_to_scene(self.scene)
        return self

    def remove_component(self, component):
        self.components.remove(component)
        if self.scene is not None:
            component._on_remove_from_scene(self.scene)
        component.entity = None

    def find_component_by_type(self, cls):
        for c in self.components:
            if isinstance(c, cls):
                return c
        return None

    def add_to_scene(self, scene):
        scene.add_entity(self)
        return self

    def remove_from_scene(self):
        if self.scene is not None:
            self.scene.remove_entity(self)


class Scene:
    """`sapien.Scene(systems=[physx_system, render_system])` (sapien_env.py:1199-1218): one sub-scene."""

    def __init__(self, systems=None):
        from . import physx as _physx
        from . import render as _render
        self.systems = list(systems) if systems is not None else [_physx.PhysxCpuSystem()]
        self.entities = []
        self.physx_system = next((s for s in self.systems if isinstance(s, _physx.PhysxSystem)), None)
        self.render_system = next((s for s in self.systems if isinstance(s, _render.RenderSystem)), None)
        self._next_id = 1
        self.ambient_light = [0.0, 0.0, 0.0]
        self.environment_map = None
        if self.physx_system is not None:
            self.physx_system._register_scene(self)
        if self.render_system is not None:
            self.render_system.scene = self

    def get_physx_system(self):
        return self.physx_system

    def get_render_system(self):
        return self.render_system

    def add_entity(self, entity):
        if entity.scene is not None:
            raise RuntimeError("entity is already in a scene")
        entity.scene = self
        entity.per_scene_id = self._next_id
        self._next_id += 1
        self.entities.append(entity)
        for c in entity.components:
            c._on_add_to_scene(self)
        return entity

    def remove_entity(self, entity):
        for c in entity.components:
            c._on_remove_from_scene(self)
        self.entities.remove(entity)
        entity.scene = None

    def get_entities(self):
        return self.entities

    def get_all_actors(self):
        from . import physx as _physx
        return [e for e in self.entities if e.find_component_by_type(_physx.PhysxRigidBaseComponent) is not None and
                e.find_component_by_type(_physx.PhysxArticulationLinkComponent) is None]

    def get_all_articulations(self):
        from . import physx as _physx
        arts = []
        for e in self.entities:
            c = e.find_component_by_type(_physx.PhysxArticulationLinkComponent)
            if c is not None and c.articulation not in arts:
                arts.append(c.articulation)
        return arts

    def update_render(self):
        """Poses reach the rasteriser straight from `cuda_rigid_body_data` when a picture is taken."""

    def step(self):
        self.physx_system.step()

    @property
    def timestep(self):
        return self.physx_system.timestep

    @timestep.setter
    def timestep(self, dt):
        self.physx_system.timestep = dt

    def set_timestep(self, dt):
        self.physx_system.timestep = dt

    def get_timestep(self):
        return self.physx_system.timestep

    # -- lighting / environment (recorded; the rasteriser uses the reference's default two directional lights + ambient)
    def set_ambient_light(self, color):
        self.ambient_light = list(color)

    def set_environment_map(self, cubemap):
        self.environment_map = cubemap

    def add_directional_light(self, direction, color, shadow=False, position=(0, 0, 0), shadow_scale=10.0, shadow_near=-10.0, shadow_far=10.0, shadow_map_size=2048):
        from . import render as _render
        e = Entity()
        light = _render.RenderDirectionalLightComponent()
        light.color, light.direction, light.shadow = list(color), list(direction), shadow
        e.add_component(light)
        self.add_entity(e)
        return light

    def add_point_light(self, position, color, shadow=False, shadow_near=0.1, shadow_far=10.0, shadow_map_size=2048):
        from . import render as _render
        e = Entity()
        light = _render.RenderPointLightComponent()
        light.color, light.shadow = list(color), shadow
        e.add_component(light)
        e.pose = Pose(position)
        self.add_entity(e)
        return light

    # -- drives between bodies (drive.py:48-50)
    def create_drive(self, body0, pose0, body1, pose1):
        from . import physx as _physx
        d = _physx.PhysxDriveComponent(body1)
        d.parent, d.pose_in_parent, d.pose_in_child = body0, pose0, pose1
        (body1.entity if body1 is not None and body1.entity is not None else Entity()).add_component(d)
        return d

    def create_actor_builder(self):
        from .wrapper.actor_builder import ActorBuilder
        return ActorBuilder().set_scene(self)

    def create_articulation_builder(self):
        from .wrapper.articulation_builder import ArticulationBuilder
        return ArticulationBuilder().set_scene(self)

    def create_urdf_loader(self):
        from .wrapper.urdf_loader import URDFLoader
        loader = URDFLoader()
        loader.set_scene(self)
        return loader


def set_log_level(level):
    pass


from . import math, physx, render, sensor, utils, wrapper  # noqa: E402,F401
from .wrapper.actor_builder import ActorBuilder  # noqa: E402,F401
from .wrapper.articulation_builder import ArticulationBuilder  # noqa: E402,F401
from .wrapper.urdf_loader import URDFLoader  # noqa: E402,F401

pysapien = _sys.modules[__name__]      # `sapien.pysapien.Entity` etc. are the same objects
core = _sys.modules[__name__]
_sys.modules.setdefault(__name__ + ".pysapien", pysapien)
_sys.modules.setdefault(__name__ + ".core", core)
