"""`sapien.physx` on the b200sim backend.

Scene-graph classes here are recorders (see the package docstring); `PhysxGpuSystem.gpu_init()` turns what was recorded into ONE
batched world (maniskill_b200/compat/compile.py -> maniskill_b200/model.py -> b2s_world_create) and from then on delegates every
`gpu_*` call / `cuda_*` buffer to maniskill_b200/physx_shim.py (one C-ABI call each).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .. import Component, Pose

# ------------------------------------------------------------------------------------------------ module-level configuration
_GPU_ENABLED = False
_CONFIG = dict(
    shape=dict(contact_offset=0.02, rest_offset=0.0),
    body=dict(solver_position_iterations=15, solver_velocity_iterations=1, sleep_threshold=0.005),
    scene=dict(gravity=np.array([0.0, 0.0, -9.81]), bounce_threshold=2.0, enable_pcm=True, enable_tgs=True, enable_ccd=False,
               enable_enhanced_determinism=False, enable_friction_every_iteration=True, cpu_workers=0),
    material=dict(static_friction=0.3, dynamic_friction=0.3, restitution=0.0),
    gpu_memory=dict(),
)


def enable_gpu():
    global _GPU_ENABLED
    _GPU_ENABLED = True


def is_gpu_enabled() -> bool:
    return _GPU_ENABLED


def set_gpu_memory_config(**kw):
    """PhysX's global buffer capacities (types.py:16-32) have no counterpart: capacities here are per sub-scene (`max_contacts`,
    `max_manifolds` of the compiled model) and overflows are reported by `World.check_overflow`."""
    _CONFIG["gpu_memory"].update(kw)


def set_shape_config(contact_offset=None, rest_offset=None):
    if contact_offset is not None:
        _CONFIG["shape"]["contact_offset"] = float(contact_offset)
    if rest_offset is not None:
        _CONFIG["shape"]["rest_offset"] = float(rest_offset)


def set_body_config(solver_position_iterations=None, solver_velocity_iterations=None, sleep_threshold=None):
    for k, v in (("solver_position_iterations", solver_position_iterations), ("solver_velocity_iterations", solver_velocity_iterations),
                 ("sleep_threshold", sleep_threshold)):
        if v is not None:
            _CONFIG["body"][k] = v


def set_scene_config(gravity=None, bounce_threshold=None, enable_pcm=None, enable_tgs=None, enable_ccd=None, enable_enhanced_determinism=None,
                     enable_friction_every_iteration=None, cpu_workers=None):
    loc = dict(gravity=gravity, bounce_threshold=bounce_threshold, enable_pcm=enable_pcm, enable_tgs=enable_tgs, enable_ccd=enable_ccd,
               enable_enhanced_determinism=enable_enhanced_determinism, enable_friction_every_iteration=enable_friction_every_iteration, cpu_workers=cpu_workers)
    for k, v in loc.items():
        if v is not None:
            _CONFIG["scene"][k] = np.asarray(v, dtype=np.float64) if k == "gravity" else v


def set_default_material(static_friction, dynamic_friction, restitution):
    _CONFIG["material"].update(static_friction=float(static_friction), dynamic_friction=float(dynamic_friction), restitution=float(restitution))


def get_default_material():
    return PhysxMaterial(**_CONFIG["material"])


def version():
    return "b200sim"


# ------------------------------------------------------------------------------------------------ materials and shapes
class PhysxMaterial:
    def __init__(self, static_friction: float = 0.3, dynamic_friction: float = 0.3, restitution: float = 0.0):
        self.static_friction, self.dynamic_friction, self.restitution = float(static_friction), float(dynamic_friction), float(restitution)

    def get_static_friction(self):
        return self.static_friction

    def get_dynamic_friction(self):
        return self.dynamic_friction

    def get_restitution(self):
        return self.restitution

    def set_static_friction(self, v):
        self.static_friction = float(v)

    def set_dynamic_friction(self, v):
        self.dynamic_friction = float(v)

    def set_restitution(self, v):
        self.restitution = float(v)


class PhysxCollisionShape:
    kind = "none"

    def __init__(self, material: Optional[PhysxMaterial] = None):
        self.physical_material = material if material is not None else get_default_material()
        self.local_pose = Pose()
        self.collision_groups = [1, 1, 0, 0]
        self.density = 1000.0
        self.patch_radius = 0.0
        self.min_patch_radius = 0.0
        self.contact_offset = _CONFIG["shape"]["contact_offset"]
        self.rest_offset = _CONFIG["shape"]["rest_offset"]
        self.parent = None

    material = property(lambda self: self.physical_material)

    def get_physical_material(self):
        return self.physical_material

    def set_physical_material(self, m):
        self.physical_material = m

    def get_local_pose(self):
        return self.local_pose

    def set_local_pose(self, pose):
        self.local_pose = pose

    def get_collision_groups(self):
        return list(self.collision_groups)

    def set_collision_groups(self, groups):
        if len(groups) != 4:
            raise RuntimeError("collision groups are 4 x uint32")
        self.collision_groups = [int(g) & 0xFFFFFFFF for g in groups]
        _changed(self.parent)

    def set_density(self, v):
        self.density = float(v)

    def get_density(self):
        return self.density

    def set_patch_radius(self, v):
        self.patch_radius = float(v)

    def get_patch_radius(self):
        return self.patch_radius

    def set_min_patch_radius(self, v):
        self.min_patch_radius = float(v)

    def get_min_patch_radius(self):
        return self.min_patch_radius

    def set_contact_offset(self, v):
        self.contact_offset = float(v)

    def set_rest_offset(self, v):
        self.rest_offset = float(v)


class PhysxCollisionShapePlane(PhysxCollisionShape):
    kind = "plane"


class PhysxCollisionShapeBox(PhysxCollisionShape):
    kind = "box"

    def __init__(self, half_size, material=None):
        super().__init__(material)
        self.half_size = np.asarray(half_size, dtype=np.float32).reshape(3)

    def get_half_size(self):
        return self.half_size


class PhysxCollisionShapeSphere(PhysxCollisionShape):
    kind = "sphere"

    def __init__(self, radius, material=None):
        super().__init__(material)
        self.radius = float(radius)

    def get_radius(self):
        return self.radius


class PhysxCollisionShapeCapsule(PhysxCollisionShape):
    kind = "capsule"

    def __init__(self, radius, half_length, material=None):
        super().__init__(material)
        self.radius, self.half_length = float(radius), float(half_length)

    def get_radius(self):
        return self.radius

    def get_half_length(self):
        return self.half_length


class PhysxCollisionShapeCylinder(PhysxCollisionShapeCapsule):
    kind = "cylinder"


class PhysxCollisionShapeConvexMesh(PhysxCollisionShape):
    """Convex mesh: the file's points (or the given vertices) are cooked to a hull of at most 64 vertices (PhysX's GPU limit)."""
    kind = "convex_mesh"

    def __init__(self, filename: str = None, scale=(1, 1, 1), material=None, vertices=None):
        super().__init__(material)
        from maniskill_b200 import meshio
        self.filename = filename
        self.scale = np.asarray(scale, dtype=np.float32).reshape(3)
        pts = np.asarray(vertices, dtype=np.float64) if vertices is not None else meshio.load_points(filename)
        try:
            self.vertices, self.triangles = meshio.cook_hull(pts * self.scale.astype(np.float64))
        except Exception as e:  # degenerate input: sapien raises RuntimeError("failed to cook mesh"), which the builders catch
            raise RuntimeError(f"failed to cook a convex mesh from {filename}: {e}")

    def get_vertices(self):
        return np.asarray(self.vertices, dtype=np.float32)

    def get_triangles(self):
        return np.asarray(self.triangles, dtype=np.uint32)

    def get_scale(self):
        return self.scale

    @staticmethod
    def load_multiple(filename: str, scale=(1, 1, 1), material=None) -> List["PhysxCollisionShapeConvexMesh"]:
        from maniskill_b200 import meshio
        return [PhysxCollisionShapeConvexMesh(filename=filename, scale=scale, material=material, vertices=part) for part in meshio.load_parts(filename)]


class PhysxCollisionShapeTriangleMesh(PhysxCollisionShapeConvexMesh):
    """Non-convex triangle meshes (static scenery) collide through their convex hull here: fine for slabs and blocks, wrong for anything an
    object is meant to sit INSIDE of (arenas, bowls) -- a warning says so once per file."""
    _warned = set()

    def __init__(self, filename: str = None, scale=(1, 1, 1), material=None, vertices=None, triangles=None):
        super().__init__(filename=filename, scale=scale, material=material, vertices=vertices)
        key = filename or "<vertices>"
        if key not in PhysxCollisionShapeTriangleMesh._warned:
            PhysxCollisionShapeTriangleMesh._warned.add(key)
            import warnings
            warnings.warn(f"non-convex collision mesh {key}: this backend collides its convex hull ({len(self.vertices)} vertices)")


# ------------------------------------------------------------------------------------------------ components
def _changed(component):
    """A property that is baked into the compiled world was changed: legal before `gpu_init`, refused afterwards."""
    sysm = getattr(component, "_system", None) if component is not None else None
    if sysm is not None and sysm._world is not None:
        raise RuntimeError("this property is compiled into the batched world at gpu_init(); change it before the scene is set up "
                           "(or reconfigure the environment)")


class PhysxBaseComponent(Component):
    def __init__(self):
        super().__init__()
        self._system = None
        self._scene = None

    def _on_add_to_scene(self, scene):
        self._scene = scene
        self._system = scene.physx_system
        if self._system is not None:
            self._system._register_component(self, scene)

    def _on_remove_from_scene(self, scene):
        if self._system is not None:
            self._system._unregister_component(self)
        self._scene = self._system = None


class PhysxRigidBaseComponent(PhysxBaseComponent):
    def __init__(self):
        super().__init__()
        self.collision_shapes: List[PhysxCollisionShape] = []
        self._pose = Pose()
        self.gpu_index = -1
        self.gpu_pose_index = -1
        self._env = -1
        self._row = -1

    def attach(self, shape: PhysxCollisionShape):
        _changed(self)
        shape.parent = self
        self.collision_shapes.append(shape)
        return self

    def get_collision_shapes(self):
        return self.collision_shapes

    def _body_pose(self):
        return self._pose

    def _set_body_pose(self, pose):
        self._pose = pose

    @property
    def pose(self):
        return self._pose

    @pose.setter
    def pose(self, pose):
        self._pose = pose
        if self.entity is not None:
            self.entity._pose = pose

    def get_pose(self):
        return self.pose

    def set_pose(self, pose):
        self.pose = pose

    def get_gpu_index(self):
        return self.gpu_index

    def get_gpu_pose_index(self):
        return self.gpu_pose_index

    def compute_global_aabb_tight(self):
        from maniskill_b200.compat.compile import shape_world_points
        pts = np.concatenate([shape_world_points(s, self.pose) for s in self.collision_shapes] or [np.zeros((1, 3))])
        return np.stack([pts.min(0), pts.max(0)])

    def get_global_aabb_fast(self):
        return self.compute_global_aabb_tight()


class PhysxRigidStaticComponent(PhysxRigidBaseComponent):
    pass


class PhysxRigidBodyComponent(PhysxRigidBaseComponent):
    def __init__(self):
        super().__init__()
        self._mass = None
        self._inertia = None
        self._cmass_local_pose = None
        self.auto_compute_mass = True
        self.linear_damping = 0.0
        self.angular_damping = 0.05
        self.disable_gravity = False
        self.max_depenetration_velocity = 3.0
        self.max_contact_impulse = 1e30
        self.linear_velocity = np.zeros(3, dtype=np.float32)
        self.angular_velocity = np.zeros(3, dtype=np.float32)
        self.solver_position_iterations = _CONFIG["body"]["solver_position_iterations"]
        self.solver_velocity_iterations = _CONFIG["body"]["solver_velocity_iterations"]
        self.sleep_threshold = _CONFIG["body"]["sleep_threshold"]

    # mass properties: explicit when set, otherwise computed from the shapes (density) at compile time
    def _auto(self):
        from maniskill_b200.compat.compile import body_mass_props
        return body_mass_props(self)

    @property
    def mass(self):
        return self._mass if self._mass is not None else self._auto()[0]

    @mass.setter
    def mass(self, v):
        _changed(self)
        self._mass, self.auto_compute_mass = float(v), False

    @property
    def inertia(self):
        return np.asarray(self._inertia if self._inertia is not None else self._auto()[2], dtype=np.float32)

    @inertia.setter
    def inertia(self, v):
        _changed(self)
        self._inertia, self.auto_compute_mass = np.asarray(v, dtype=np.float64).reshape(3), False

    @property
    def cmass_local_pose(self):
        return self._cmass_local_pose if self._cmass_local_pose is not None else self._auto()[1]

    @cmass_local_pose.setter
    def cmass_local_pose(self, pose):
        _changed(self)
        self._cmass_local_pose, self.auto_compute_mass = pose, False

    def get_mass(self):
        return self.mass

    def set_mass(self, v):
        self.mass = v

    def get_inertia(self):
        return self.inertia

    def set_inertia(self, v):
        self.inertia = v

    def get_cmass_local_pose(self):
        return self.cmass_local_pose

    def set_cmass_local_pose(self, p):
        self.cmass_local_pose = p

    def get_auto_compute_mass(self):
        return self.auto_compute_mass

    def get_linear_velocity(self):
        return self.linear_velocity

    def get_angular_velocity(self):
        return self.angular_velocity

    def set_linear_velocity(self, v):
        self.linear_velocity = np.asarray(v, dtype=np.float32).reshape(3)

    def set_angular_velocity(self, v):
        self.angular_velocity = np.asarray(v, dtype=np.float32).reshape(3)

    def get_linear_damping(self):
        return self.linear_damping

    def set_linear_damping(self, v):
        _changed(self)
        self.linear_damping = float(v)

    def get_angular_damping(self):
        return self.angular_damping

    def set_angular_damping(self, v):
        _changed(self)
        self.angular_damping = float(v)

    def get_disable_gravity(self):
        return self.disable_gravity

    def set_disable_gravity(self, v):
        _changed(self)
        self.disable_gravity = bool(v)

    def set_max_depenetration_velocity(self, v):
        self.max_depenetration_velocity = float(v)

    def set_max_contact_impulse(self, v):
        self.max_contact_impulse = float(v)

    def add_force_at_point(self, force, point, mode="force"):
        raise NotImplementedError("external forces on single bodies go through the batched buffers on the GPU backend")

    def add_force_torque(self, force, torque, mode="force"):
        raise NotImplementedError("external forces on single bodies go through the batched buffers on the GPU backend")


class PhysxRigidDynamicComponent(PhysxRigidBodyComponent):
    def __init__(self):
        super().__init__()
        self._kinematic = False
        self.locked_motion_axes = [False] * 6
        self.is_sleeping = False
        self.kinematic_target = None

    @property
    def kinematic(self):
        return self._kinematic

    @kinematic.setter
    def kinematic(self, v):
        _changed(self)
        self._kinematic = bool(v)

    def get_kinematic(self):
        return self._kinematic

    def set_kinematic(self, v):
        self.kinematic = v

    def get_locked_motion_axes(self):
        return list(self.locked_motion_axes)

    def set_locked_motion_axes(self, axes):
        if any(axes):
            raise NotImplementedError("locked motion axes are not supported by the b200sim backend")
        self.locked_motion_axes = [bool(a) for a in axes]

    def wake_up(self):
        pass

    def put_to_sleep(self):
        pass

    def set_kinematic_target(self, pose):
        self.kinematic_target = pose


class PhysxArticulationJoint:
    """One joint of an articulation = the incoming joint of a link (joint frame: x axis = motion axis)."""

    def __init__(self, child_link, parent_link):
        self.child_link = child_link
        self.parent_link = parent_link
        self.name = ""
        self._type = "fixed" if parent_link is not None else "undefined"
        self.pose_in_parent = Pose()
        self.pose_in_child = Pose()
        self._limit = np.array([[-np.inf, np.inf]], dtype=np.float32)
        self.friction = 0.0
        self.armature = np.zeros(0, dtype=np.float32)
        self.stiffness = 0.0
        self.damping = 0.0
        self.force_limit = 3.4028234663852886e38
        self.drive_mode = "force"
        self.drive_target = np.zeros(1, dtype=np.float32)
        self.drive_velocity_target = np.zeros(1, dtype=np.float32)

    @property
    def type(self):
        return self._type

    @type.setter
    def type(self, t):
        if t not in ("fixed", "revolute", "revolute_unwrapped", "continuous", "prismatic", "free", "undefined", "spherical"):
            raise RuntimeError(f"invalid joint type {t}")
        _changed(self.child_link)
        self._type = t

    def get_type(self):
        return self._type

    def set_type(self, t):
        self.type = t

    @property
    def dof(self):
        return {"revolute": 1, "revolute_unwrapped": 1, "continuous": 1, "prismatic": 1, "spherical": 3, "free": 6}.get(self._type, 0)

    def get_dof(self):
        return self.dof

    @property
    def limit(self):
        return self._limit if self.dof else np.zeros((0, 2), dtype=np.float32)

    @limit.setter
    def limit(self, v):
        _changed(self.child_link)
        self._limit = np.asarray(v, dtype=np.float32).reshape(-1, 2)

    limits = limit

    def get_limit(self):
        return self.limit

    def set_limit(self, v):
        self.limit = v

    get_limits, set_limits = get_limit, set_limit

    def get_name(self):
        return self.name

    def set_name(self, n):
        self.name = n

    def get_parent_link(self):
        return self.parent_link

    def get_child_link(self):
        return self.child_link

    def get_pose_in_parent(self):
        return self.pose_in_parent

    def get_pose_in_child(self):
        return self.pose_in_child

    def set_pose_in_parent(self, p):
        self.pose_in_parent = p

    def set_pose_in_child(self, p):
        self.pose_in_child = p

    def get_global_pose(self):
        return self.child_link.pose * self.pose_in_child

    # drive: tau = stiffness (target - q) + damping (target_velocity - qd), clamped to +-force_limit (articulation_joint.py:187-195)
    def set_drive_properties(self, stiffness, damping, force_limit=3.4028234663852886e38, mode="force"):
        if mode not in ("force", "acceleration"):
            raise RuntimeError(f"invalid drive mode {mode}")
        sysm = getattr(self.child_link, "_system", None)
        if sysm is not None and sysm._world is not None:
            sysm._update_drive(self, float(stiffness), float(damping), float(force_limit), mode)
        self.stiffness, self.damping, self.force_limit, self.drive_mode = float(stiffness), float(damping), float(force_limit), mode

    def set_drive_property(self, stiffness, damping, force_limit=3.4028234663852886e38, mode="force"):
        self.set_drive_properties(stiffness, damping, force_limit, mode)

    def get_stiffness(self):
        return self.stiffness

    def get_damping(self):
        return self.damping

    def get_force_limit(self):
        return self.force_limit

    def get_drive_mode(self):
        return self.drive_mode

    def get_friction(self):
        return self.friction

    def set_friction(self, v):
        _changed(self.child_link)
        self.friction = float(v)

    def get_armature(self):
        return self.armature

    def set_armature(self, v):
        _changed(self.child_link)
        self.armature = np.asarray(v, dtype=np.float32).reshape(-1)

    def get_drive_target(self):
        return self.drive_target

    def set_drive_target(self, v):
        self.drive_target = np.asarray(v, dtype=np.float32).reshape(-1)

    def get_drive_velocity_target(self):
        return self.drive_velocity_target

    def set_drive_velocity_target(self, v):
        self.drive_velocity_target = np.asarray(v, dtype=np.float32).reshape(-1)


class PhysxArticulationLinkComponent(PhysxRigidBodyComponent):
    def __init__(self, parent: Optional["PhysxArticulationLinkComponent"] = None):
        super().__init__()
        self.parent = parent
        self.children: List[PhysxArticulationLinkComponent] = []
        if parent is None:
            self.articulation = PhysxArticulation()
        else:
            self.articulation = parent.articulation
            parent.children.append(self)
        self.index = len(self.articulation.links)
        self.articulation.links.append(self)
        self.joint = PhysxArticulationJoint(self, parent)
        self.sleeping = False

    @property
    def is_root(self):
        return self.parent is None

    def get_parent(self):
        return self.parent

    def get_children(self):
        return self.children

    def get_articulation(self):
        return self.articulation

    def get_index(self):
        return self.index

    def get_joint(self):
        return self.joint

    def put_to_sleep(self):
        pass

    def wake_up(self):
        pass

    def _set_body_pose(self, pose):
        self._pose = pose
        if self.parent is None:
            self.articulation._root_pose = pose

    @property
    def pose(self):
        return self._pose

    @pose.setter
    def pose(self, pose):   # the root link's pose IS the articulation's root pose (mani_skill/utils/structs/articulation.py:857-859 sets it through the link)
        PhysxRigidBaseComponent.pose.fset(self, pose)
        if self.parent is None:
            self.articulation._root_pose = pose


class PhysxArticulationLink(PhysxArticulationLinkComponent):
    pass


class PhysxArticulation:
    def __init__(self):
        self.links: List[PhysxArticulationLinkComponent] = []
        self.name = ""
        self._root_pose = Pose()
        self.gpu_index = -1
        self.tendons = []
        self.srdf_disabled_pairs = []   # link-name pairs the SRDF lists under <disable_collisions>
        self._qpos = None
        self._env = -1

    @property
    def root(self):
        return self.links[0]

    def get_root(self):
        return self.links[0]

    def get_links(self):
        return self.links

    @property
    def joints(self):
        return [l.joint for l in self.links]

    def get_joints(self):
        return self.joints

    @property
    def active_joints(self):
        return [l.joint for l in self.links if l.joint.dof > 0]

    def get_active_joints(self):
        return self.active_joints

    @property
    def dof(self):
        return sum(l.joint.dof for l in self.links)

    def get_dof(self):
        return self.dof

    def get_name(self):
        return self.name

    def set_name(self, n):
        self.name = n

    def get_gpu_index(self):
        return self.gpu_index

    @property
    def pose(self):
        return self._root_pose

    @pose.setter
    def pose(self, p):
        _changed(self.links[0] if self.links else None)
        self._root_pose = p
        if self.links:
            self.links[0]._pose = p
            if self.links[0].entity is not None:
                self.links[0].entity._pose = p

    root_pose = pose

    def get_pose(self):
        return self._root_pose

    def set_pose(self, p):
        self.pose = p

    get_root_pose, set_root_pose = get_pose, set_pose

    def find_link_by_name(self, name):
        return next((l for l in self.links if l.name == name), None)

    def find_joint_by_name(self, name):
        return next((l.joint for l in self.links if l.joint.name == name), None)

    # generalized coordinates of the un-simulated prototype (initial values); live values are in the cuda_* buffers
    def _vec(self, v):
        return np.zeros(self.dof, dtype=np.float32) if v is None else v

    @property
    def qpos(self):
        return self._vec(self._qpos)

    @qpos.setter
    def qpos(self, v):
        self._qpos = np.asarray(v, dtype=np.float32).reshape(-1)

    qvel = qacc = qf = property(lambda self: np.zeros(self.dof, dtype=np.float32), lambda self, v: None)

    def get_qpos(self):
        return self.qpos

    def set_qpos(self, v):
        self.qpos = v

    def get_qvel(self):
        return np.zeros(self.dof, dtype=np.float32)

    def set_qvel(self, v):
        pass

    def get_qf(self):
        return np.zeros(self.dof, dtype=np.float32)

    def set_qf(self, v):
        pass

    def get_qacc(self):
        return np.zeros(self.dof, dtype=np.float32)

    @property
    def qlimit(self):
        return np.concatenate([j.limit for j in self.active_joints] or [np.zeros((0, 2), dtype=np.float32)]).astype(np.float32)

    qlimits = qlimit

    def get_qlimit(self):
        return self.qlimit

    get_qlimits = get_qlimit

    def set_root_linear_velocity(self, v):
        pass

    def set_root_angular_velocity(self, v):
        pass

    set_root_velocity = set_root_linear_velocity

    def get_root_linear_velocity(self):
        return np.zeros(3, dtype=np.float32)

    get_root_angular_velocity = get_root_velocity = get_root_linear_velocity

    def create_fixed_tendon(self, link_chain, coefficients, recip_coefficients, rest_length=0, offset=0, stiffness=0, damping=0, low=-3.4028234663852886e38,
                            high=3.4028234663852886e38, limit_stiffness=0):
        """articulation_builder.py:161-200: the mimic coupling of two joints, recorded as q_child = -c1/c2 ... (see compile.py)."""
        _changed(self.links[0])
        self.tendons.append(dict(chain=list(link_chain), coefficients=list(coefficients), recip=list(recip_coefficients), rest_length=float(rest_length),
                                 offset=float(offset), stiffness=float(stiffness), damping=float(damping)))

    def compute_passive_force(self, gravity=True, coriolis_and_centrifugal=True):
        raise NotImplementedError("PhysxArticulation.compute_passive_force is a CPU-simulation call")

    def get_link_incoming_joint_forces(self):
        raise NotImplementedError("link incoming joint forces are not computed by the b200sim backend")

    def create_pinocchio_model(self):
        raise NotImplementedError("pinocchio models belong to the CPU simulation path")


class PhysxDriveComponent(PhysxBaseComponent):
    """6-D drive between two bodies (drive.py:48-50).  Recorded only: the b200sim solver has no body-to-body drives."""

    def __init__(self, body=None):
        super().__init__()
        self.child = body
        self.parent = None
        self.pose_in_parent = self.pose_in_child = Pose()

    def _unsupported(self, *a, **kw):
        raise NotImplementedError("PhysxDriveComponent constraints are not supported by the b200sim backend")

    set_drive_property_x = set_drive_property_y = set_drive_property_z = set_limit_x = set_limit_y = set_limit_z = _unsupported
    set_drive_property_twist = set_drive_property_swing = set_drive_property_slerp = set_limit_twist = set_limit_cone = _unsupported


class PhysxJointComponent(PhysxDriveComponent):
    pass


class PhysxGearComponent(PhysxDriveComponent):
    pass


class PhysxDistanceJointComponent(PhysxDriveComponent):
    pass


class PhysxContactPoint:
    __slots__ = ("impulse", "normal", "position", "separation")


class PhysxContact:
    __slots__ = ("bodies", "shapes", "points")


# ------------------------------------------------------------------------------------------------ systems
class PhysxSystem:
    def __init__(self):
        self.scenes = []
        self._timestep = 0.01
        self._world = None

    def _register_scene(self, scene):
        self.scenes.append(scene)

    def _register_component(self, component, scene):
        pass

    def _unregister_component(self, component):
        pass

    @property
    def timestep(self):
        return self._timestep

    @timestep.setter
    def timestep(self, dt):
        self._timestep = float(dt)

    def get_timestep(self):
        return self._timestep

    def set_timestep(self, dt):
        self.timestep = dt

    def get_config(self):
        return _CONFIG


class PhysxCpuSystem(PhysxSystem):
    def __init__(self, *a, **kw):
        raise RuntimeError("b200sim has no CPU simulation path: use sim_backend='physx_cuda' (num_envs >= 1 on a B200)")


from maniskill_b200.physx_shim import ContactImpulseQuery as PhysxGpuContactQuery  # noqa: E402

PhysxGpuContactPairImpulseQuery = PhysxGpuContactBodyImpulseQuery = PhysxGpuContactQuery


class _LazyCudaArray:
    """`px.cuda_*` before and after gpu_init: `.torch()` returns the world's aliasing tensor."""

    def __init__(self, system, name):
        self._system, self._name = system, name

    def torch(self):
        return getattr(self._system._facade, self._name).torch()

    @property
    def shape(self):
        return tuple(self.torch().shape)


class PhysxGpuSystem(PhysxSystem):
    """One batched world for all sub-scenes (sapien_env.py:1186-1210 creates ONE PhysxGpuSystem and N `sapien.Scene`s on it)."""

    _BUFFERS = ("cuda_rigid_body_data", "cuda_articulation_qpos", "cuda_articulation_qvel", "cuda_articulation_qacc", "cuda_articulation_qf",
                "cuda_articulation_target_qpos", "cuda_articulation_target_qvel")

    def __init__(self, device="cuda"):
        super().__init__()
        from .. import Device
        self.device = device if isinstance(device, Device) else Device(str(device))
        self._offsets = {}
        self.rigid_dynamic_components: List[PhysxRigidDynamicComponent] = []
        self.rigid_static_components: List[PhysxRigidStaticComponent] = []
        self.articulation_link_components: List[PhysxArticulationLinkComponent] = []
        self._by_scene = {}
        self._facade = None
        self._compiled = None
        for b in self._BUFFERS:
            setattr(self, b, _LazyCudaArray(self, b))

    # ---- scene bookkeeping
    def set_scene_offset(self, scene, offset):
        self._offsets[id(scene)] = np.asarray(offset, dtype=np.float32).reshape(3)

    def get_scene_offset(self, scene):
        return self._offsets.get(id(scene), np.zeros(3, dtype=np.float32))

    def _register_component(self, component, scene):
        if self._world is not None:
            raise RuntimeError("entities cannot be added after gpu_init(): the batched world is compiled once (reconfigure the environment instead)")
        if isinstance(component, PhysxArticulationLinkComponent):
            self.articulation_link_components.append(component)
        elif isinstance(component, PhysxRigidDynamicComponent):
            self.rigid_dynamic_components.append(component)
        elif isinstance(component, PhysxRigidStaticComponent):
            self.rigid_static_components.append(component)
        else:
            return
        self._by_scene.setdefault(id(scene), []).append(component)

    def _unregister_component(self, component):
        if self._world is not None:
            raise RuntimeError("entities cannot be removed after gpu_init() (hide them or reconfigure the environment)")
        for lst in (self.articulation_link_components, self.rigid_dynamic_components, self.rigid_static_components):
            if component in lst:
                lst.remove(component)
        for lst in self._by_scene.values():
            if component in lst:
                lst.remove(component)

    get_rigid_dynamic_components = lambda self: self.rigid_dynamic_components
    get_rigid_static_components = lambda self: self.rigid_static_components
    get_articulation_link_components = lambda self: self.articulation_link_components

    # ---- the compile step
    def gpu_init(self):
        from maniskill_b200.compat import compile as _c
        if self._world is not None:
            return
        self._compiled = _c.compile_system(self, _CONFIG)
        self._world = self._compiled.world
        self._facade = self._compiled.facade

    @PhysxSystem.timestep.setter
    def timestep(self, dt):
        if self._world is not None and abs(float(dt) - self._timestep) > 1e-12:
            raise RuntimeError(f"the batched world was compiled with timestep {self._timestep}; reconfigure to change it")
        self._timestep = float(dt)

    def _update_drive(self, joint, stiffness, damping, force_limit, mode):
        """Drive gains changed after gpu_init (e.g. a control mode switch): allowed when every sub-scene gets the same values."""
        self._compiled.update_drive(joint, stiffness, damping, force_limit, mode)

    def gpu_set_cuda_stream(self, stream):
        pass

    def sync_poses_gpu_to_cpu(self):
        self._compiled.sync_poses_to_objects()

    def step(self):
        self._facade.step()

    def get_contacts(self):
        raise NotImplementedError("per-contact lists are a CPU-simulation API; use the contact impulse queries on the GPU backend")

    def __getattr__(self, name):
        # every gpu_apply_* / gpu_fetch_* / gpu_update_* / gpu_create_* / gpu_query_* entry point is the facade's
        if name.startswith("gpu_") and self.__dict__.get("_facade") is not None:
            return getattr(self.__dict__["_facade"], name)
        if name.startswith("gpu_"):
            raise RuntimeError(f"PhysxGpuSystem.{name}() called before gpu_init()")
        raise AttributeError(name)

    # buffers this backend does not produce
    @property
    def cuda_rigid_body_force(self):
        raise NotImplementedError("cuda_rigid_body_force (external forces on free bodies) is not supported by the b200sim backend")

    @property
    def cuda_articulation_link_incoming_joint_forces(self):
        raise NotImplementedError("link incoming joint forces are not computed by the b200sim backend")
