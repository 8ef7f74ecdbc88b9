"""sapien.wrapper.urdf_loader.URDFLoader: URDF (+ SRDF) -> articulation / actor builders (the reference subclasses it,
mani_skill/utils/building/urdf_loader.py:23-47, and configures it through mani_skill/utils/sapien_utils.py:147-172).

`parse()` walks the kinematic tree from the root link in depth-first URDF joint order and fills link builders obtained from
`scene.create_articulation_builder()` -- so the scene's (ManiSkill's) builder subclasses are the ones that get built:
  * inertial -> `set_mass_and_inertia` (mass, centre-of-mass frame = principal axes, principal moments)
  * <collision> -> box / sphere / cylinder / capsule / mesh records (link materials / patch radii / densities from the loader's tables)
  * <visual>    -> the matching visual records
  * <joint>     -> joint frames with the motion axis on x (`pose_in_parent` = origin * R(axis), `pose_in_child` = R(axis)), limits,
                   friction / damping, <mimic> -> MimicJointRecord
  * SRDF <disable_collisions> pairs are recorded on the root link builder (consumed when the articulation is compiled)
Stand-alone parts of the file without joints become actor builders.  Cameras (<sensor> elements of sapien's URDF extension) are not read.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Tuple

import numpy as np

import sapien
from sapien import physx

from .articulation_builder import MimicJointRecord


def _rpy_to_mat(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _pose_from(xyz, R) -> sapien.Pose:
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, xyz
    return sapien.Pose(T)


def _origin(el) -> Tuple[np.ndarray, np.ndarray]:
    xyz, rpy = np.zeros(3), np.zeros(3)
    o = el.find("origin") if el is not None else None
    if o is not None:
        if o.get("xyz"):
            xyz = np.array([float(v) for v in o.get("xyz").split()])
        if o.get("rpy"):
            rpy = np.array([float(v) for v in o.get("rpy").split()])
    return xyz, _rpy_to_mat(*rpy)


def _axis_rotation(axis) -> np.ndarray:
    """Rotation whose first column is `axis` (the joint frame's x axis is the motion axis)."""
    x = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(x)
    x = x / n if n > 1e-12 else np.array([1.0, 0, 0])
    helper = np.array([0.0, 0, 1]) if abs(x[2]) < 0.9 else np.array([0.0, 1, 0])
    y = np.cross(helper, x)
    y /= np.linalg.norm(y)
    return np.stack([x, y, np.cross(x, y)], axis=1)


class URDFLoader:
    def __init__(self):
        self.fix_root_link = True
        self.load_multiple_collisions_from_file = False
        self.load_nonconvex_collisions_from_file = False
        self.multiple_collisions_decomposition = "none"
        self.multiple_collisions_decomposition_params = dict()
        self.revolute_unwrapped = False
        self.scale = 1.0
        self.scene = None
        self._material = None
        self._patch_radius = 0.0
        self._min_patch_radius = 0.0
        self.density = 1000.0
        self._link_material: Dict[str, physx.PhysxMaterial] = dict()
        self._link_patch_radius: Dict[str, float] = dict()
        self._link_min_patch_radius: Dict[str, float] = dict()
        self._link_density: Dict[str, float] = dict()
        self.collision_is_visual = False
        self.package_dir = None

    # ---- configuration (sapien_utils.apply_urdf_config)
    def set_scene(self, scene):
        self.scene = scene
        return self

    def set_material(self, static_friction, dynamic_friction, restitution):
        self._material = physx.PhysxMaterial(static_friction, dynamic_friction, restitution)

    def set_patch_radius(self, v):
        self._patch_radius = float(v)

    def set_min_patch_radius(self, v):
        self._min_patch_radius = float(v)

    def set_density(self, v):
        self.density = float(v)

    def set_link_material(self, link_name, static_friction, dynamic_friction, restitution):
        self._link_material[link_name] = physx.PhysxMaterial(static_friction, dynamic_friction, restitution)

    def set_link_patch_radius(self, link_name, v):
        self._link_patch_radius[link_name] = float(v)

    def set_link_min_patch_radius(self, link_name, v):
        self._link_min_patch_radius[link_name] = float(v)

    def set_link_density(self, link_name, v):
        self._link_density[link_name] = float(v)

    # ---- helpers
    def _resolve(self, filename, urdf_dir):
        if filename.startswith("package://"):
            rel = filename[len("package://"):]
            return os.path.join(self.package_dir if self.package_dir else urdf_dir, rel)
        return filename if os.path.isabs(filename) else os.path.join(urdf_dir, filename)

    def _fill_link(self, builder, link_el, urdf_dir):
        name = link_el.get("name")
        builder.set_name(name)
        s = float(self.scale)
        material = self._link_material.get(name, self._material)
        patch = self._link_patch_radius.get(name, self._patch_radius)
        min_patch = self._link_min_patch_radius.get(name, self._min_patch_radius)
        density = self._link_density.get(name, self.density)
        inertial = link_el.find("inertial")
        if inertial is not None and inertial.find("mass") is not None and float(inertial.find("mass").get("value")) > 0:
            mass = float(inertial.find("mass").get("value"))
            xyz, R = _origin(inertial)
            ie = inertial.find("inertia")
            g = (lambda k: float(ie.get(k, 0.0))) if ie is not None else (lambda k: 0.0)
            I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            w, V = np.linalg.eigh(I)
            if np.linalg.det(V) < 0:
                V[:, 2] *= -1
            builder.set_mass_and_inertia(mass * s**3, _pose_from(xyz * s, R @ V), np.maximum(w, 0.0) * s**5)
        for tag, is_col in (("collision", True), ("visual", False)):
            for el in link_el.findall(tag):
                xyz, R = _origin(el)
                pose = _pose_from(xyz * s, R)
                geo = el.find("geometry")
                if geo is None:
                    continue
                vis_material = None
                if not is_col:
                    m = el.find("material")
                    c = m.find("color") if m is not None else None
                    if c is not None and c.get("rgba"):
                        vis_material = [float(v) for v in c.get("rgba").split()]
                if geo.find("box") is not None:
                    half = [float(v) * s / 2 for v in geo.find("box").get("size").split()]
                    if is_col:
                        builder.add_box_collision(pose, half, material=material, density=density, patch_radius=patch, min_patch_radius=min_patch)
                    else:
                        builder.add_box_visual(pose, half, material=vis_material, name=el.get("name", ""))
                elif geo.find("sphere") is not None:
                    r = float(geo.find("sphere").get("radius")) * s
                    if is_col:
                        builder.add_sphere_collision(pose, r, material=material, density=density, patch_radius=patch, min_patch_radius=min_patch)
                    else:
                        builder.add_sphere_visual(pose, r, material=vis_material, name=el.get("name", ""))
                elif geo.find("cylinder") is not None or geo.find("capsule") is not None:
                    cyl = geo.find("cylinder") if geo.find("cylinder") is not None else geo.find("capsule")
                    r, hl = float(cyl.get("radius")) * s, float(cyl.get("length")) * s / 2
                    # URDF cylinders run along z, sapien's along x
                    zpose = pose * sapien.Pose(q=[0.7071068, 0, -0.7071068, 0])
                    kind = "cylinder" if geo.find("cylinder") is not None else "capsule"
                    if is_col:
                        getattr(builder, f"add_{kind}_collision")(zpose, r, hl, material=material, density=density, patch_radius=patch, min_patch_radius=min_patch)
                    else:
                        getattr(builder, f"add_{kind}_visual")(zpose, r, hl, material=vis_material, name=el.get("name", ""))
                elif geo.find("mesh") is not None:
                    m = geo.find("mesh")
                    scale = np.array([float(v) for v in m.get("scale", "1 1 1").split()]) * s
                    filename = self._resolve(m.get("filename"), urdf_dir)
                    if is_col:
                        if self.load_multiple_collisions_from_file:
                            builder.add_multiple_convex_collisions_from_file(filename, pose, scale, material=material, density=density, patch_radius=patch,
                                                                             min_patch_radius=min_patch, decomposition=self.multiple_collisions_decomposition,
                                                                             decomposition_params=self.multiple_collisions_decomposition_params)
                        elif self.load_nonconvex_collisions_from_file:
                            builder.add_nonconvex_collision_from_file(filename, pose, scale, material=material, patch_radius=patch, min_patch_radius=min_patch)
                        else:
                            builder.add_convex_collision_from_file(filename, pose, scale, material=material, density=density, patch_radius=patch,
                                                                   min_patch_radius=min_patch)
                    else:
                        builder.add_visual_from_file(filename, pose, scale, material=vis_material, name=el.get("name", ""))

    # ---- parse
    def parse(self, urdf_file, srdf_file=None, package_dir=None):
        """-> (articulation_builders, actor_builders, cameras)."""
        self.package_dir = package_dir
        urdf_file = str(urdf_file)
        urdf_dir = os.path.dirname(os.path.abspath(urdf_file))
        root = ET.parse(urdf_file).getroot()
        links = {l.get("name"): l for l in root.findall("link")}
        children: Dict[str, List[ET.Element]] = {}
        child_names = set()
        for j in root.findall("joint"):
            children.setdefault(j.find("parent").get("link"), []).append(j)
            child_names.add(j.find("child").get("link"))
        roots = [n for n in links if n not in child_names]
        if srdf_file is None and os.path.exists(urdf_file[:-4] + "srdf"):
            srdf_file = urdf_file[:-4] + "srdf"
        disabled = []
        if srdf_file is not None and os.path.exists(str(srdf_file)):
            for d in ET.parse(str(srdf_file)).getroot().findall("disable_collisions"):
                disabled.append((d.get("link1"), d.get("link2")))
        articulation_builders, actor_builders = [], []
        s = float(self.scale)
        for rname in roots:
            if rname not in children:  # a link without joints: a plain actor
                b = self.scene.create_actor_builder()
                self._fill_link(b, links[rname], urdf_dir)
                actor_builders.append(b)
                continue
            ab = self.scene.create_articulation_builder()
            mimics = []

            seen = set()

            def visit(lname, joint_el, parent_builder):
                if lname in seen or lname not in links:
                    raise RuntimeError(f"URDF {urdf_file}: link '{lname}' " + ("is its own ancestor (joint cycle)" if lname in seen else "is referenced by a joint but not defined"))
                seen.add(lname)
                lb = ab.create_link_builder(parent_builder)
                self._fill_link(lb, links[lname], urdf_dir)
                if joint_el is None:
                    lb.set_joint_name("")
                    lb.set_joint_properties("fixed" if self.fix_root_link else "undefined", [], sapien.Pose(), sapien.Pose(), 0, 0)
                else:
                    jt = joint_el.get("type")
                    xyz, R = _origin(joint_el)
                    axis = [1.0, 0.0, 0.0]
                    if joint_el.find("axis") is not None:
                        axis = [float(v) for v in joint_el.find("axis").get("xyz").split()]
                    Ra = _axis_rotation(axis)
                    lim = joint_el.find("limit")
                    lo = float(lim.get("lower", 0)) if lim is not None else 0.0
                    hi = float(lim.get("upper", 0)) if lim is not None else 0.0
                    dyn = joint_el.find("dynamics")
                    damping = float(dyn.get("damping", 0)) if dyn is not None else 0.0
                    friction = float(dyn.get("friction", 0)) if dyn is not None else 0.0
                    limits = []
                    if jt == "revolute":
                        jtype, limits = ("revolute_unwrapped" if self.revolute_unwrapped else "revolute"), [[lo, hi]]
                    elif jt == "continuous":
                        jtype, limits = "revolute_unwrapped", [[-np.inf, np.inf]]
                    elif jt == "prismatic":
                        jtype, limits = "prismatic", [[lo * s, hi * s]]
                    elif jt == "fixed":
                        jtype = "fixed"
                    elif jt == "floating":
                        jtype = "free"
                    else:
                        raise RuntimeError(f"unsupported URDF joint type '{jt}' ({joint_el.get('name')})")
                    lb.set_joint_name(joint_el.get("name"))
                    lb.set_joint_properties(jtype, limits, _pose_from(xyz * s, R @ Ra), _pose_from(np.zeros(3), Ra), friction, damping)
                    mm = joint_el.find("mimic")
                    if mm is not None:
                        mimics.append(MimicJointRecord(joint_el.get("name"), mm.get("joint"), float(mm.get("multiplier", 1)), float(mm.get("offset", 0))))
                for j in children.get(lname, []):
                    visit(j.find("child").get("link"), j, lb)
                return lb

            root_builder = visit(rname, None, None)
            ab.mimic_joint_records = mimics
            root_builder._srdf_disabled = disabled
            articulation_builders.append(ab)
        return articulation_builders, actor_builders, []

    def load_file_as_articulation_builder(self, urdf_file, srdf_file=None, package_dir=None):
        arts, actors, cams = URDFLoader.parse(self, urdf_file, srdf_file, package_dir)
        if len(arts) != 1 or actors:
            raise Exception("URDF contains multiple objects, call load_multiple instead")
        return arts[0]

    def load(self, urdf_file, srdf_file=None, package_dir=None):
        ab = self.load_file_as_articulation_builder(urdf_file, srdf_file, package_dir)
        return ab.build(fix_root_link=self.fix_root_link)
