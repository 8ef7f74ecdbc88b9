"""Scene description -> compiled model tables (the ``B2SModel`` struct of include/b200sim.h).

Host-side counterpart of what ManiSkill hands to sapien while building a scene:
``ActorBuilder`` records (mani_skill/utils/building/actor_builder.py:57-164), ``ArticulationBuilder`` / URDF loader
link + joint records (mani_skill/utils/building/articulation_builder.py:65-112, urdf_loader.py:23-123), collision
groups (actor_builder.py:151; rule documented at docs/source/user_guide/tutorials/custom_robots.md:453) and SRDF
``disable_collisions`` pairs.  ``compile()`` turns ONE sub-scene prototype into flat arrays; the backend
instantiates it ``n_envs`` times on the device (struct-of-arrays), instead of the reference's N python object graphs
(mani_skill/envs/sapien_env.py:1186-1210).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

SHAPE_PLANE, SHAPE_BOX, SHAPE_SPHERE, SHAPE_CAPSULE, SHAPE_CONVEX = 0, 1, 2, 3, 4
OWNER_STATIC, OWNER_LINK, OWNER_BODY = 0, 1, 2
JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1
BODY_DYNAMIC, BODY_KINEMATIC = 0, 1
DEFAULT_DENSITY = 1000.0  # sapien's default shape density (not in the reference tree)


# ----------------------------------------------------------------------------- small pose algebra (numpy, float64)
def qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def qrot(q, v):
    u = np.asarray(q[1:], dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    t = 2 * np.cross(u, v)
    return v + q[0] * t + np.cross(u, t)


def qmat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def pose_mul(a, b):
    """a, b: 7-vectors (p, q wxyz)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.empty(7)
    out[:3] = a[:3] + qrot(a[3:], b[:3])
    q = qmul(a[3:], b[3:])
    out[3:] = q / np.linalg.norm(q)
    return out


def pose_inv(a):
    a = np.asarray(a, dtype=np.float64)
    qi = np.array([a[3], -a[4], -a[5], -a[6]])
    out = np.empty(7)
    out[:3] = -qrot(qi, a[:3])
    out[3:] = qi
    return out


IDENTITY = np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


def pose7(p=(0, 0, 0), q=(1, 0, 0, 0)):
    return np.concatenate([np.asarray(p, dtype=np.float64), np.asarray(q, dtype=np.float64)])


# ----------------------------------------------------------------------------- mass properties
def box_mass(half, density):
    hx, hy, hz = half
    m = density * 8 * hx * hy * hz
    I = np.diag([m / 3 * (hy * hy + hz * hz), m / 3 * (hx * hx + hz * hz), m / 3 * (hx * hx + hy * hy)])
    return m, np.zeros(3), I


def sphere_mass(r, density):
    m = density * 4.0 / 3.0 * np.pi * r**3
    return m, np.zeros(3), np.eye(3) * (0.4 * m * r * r)


def capsule_mass(r, hl, density):
    # cylinder (axis x) + two hemispheres
    mc = density * np.pi * r * r * 2 * hl
    ms = density * 4.0 / 3.0 * np.pi * r**3
    m = mc + ms
    ixx = 0.5 * mc * r * r + 0.4 * ms * r * r
    iyy = mc * (r * r / 4 + (2 * hl) ** 2 / 12) + ms * (0.4 * r * r + hl * hl + 0.75 * r * hl)
    return m, np.zeros(3), np.diag([ixx, iyy, iyy])


def hull_mass(verts, tris, density):
    """Exact polyhedral mass properties (signed tetrahedra against the origin)."""
    v = np.asarray(verts, dtype=np.float64)
    vol = 0.0
    com = np.zeros(3)
    Cc = np.zeros((3, 3))
    canon = np.array([[2, 1, 1], [1, 2, 1], [1, 1, 2]]) / 120.0
    for t in tris:
        a, b, c = v[t[0]], v[t[1]], v[t[2]]
        A = np.stack([a, b, c], 1)
        d = np.linalg.det(A)
        vol += d / 6
        com += d / 24 * (a + b + c)
        Cc += d * A @ canon @ A.T
    com /= vol
    Cc = Cc - vol * np.outer(com, com)  # covariance about com
    I = density * (np.trace(Cc) * np.eye(3) - Cc)
    return density * vol, com, I


def cylinder_shape(radius, half_length, pose=None, segments=24, **kw):
    """`add_cylinder_collision(radius, half_length)` (mani_skill/utils/building/actor_builder.py:104-116; axis = local x): like
    SAPIEN/PhysX, a cylinder is a cooked convex mesh -- here a `segments`-sided prism handled by the convex paths (plane: vertex
    tests, box: GJK/EPA + vertex patch); mass properties come from the polyhedron.  Returns a ShapeRec (keyword arguments are
    passed on: mu, density, groups, color ...)."""
    from scipy.spatial import ConvexHull
    ang = 2 * np.pi * (np.arange(segments) + 0.5) / segments
    ring = np.stack([np.cos(ang), np.sin(ang)], 1) * radius
    verts = np.concatenate([np.concatenate([np.full((segments, 1), sx * half_length), ring], 1) for sx in (-1.0, 1.0)])
    hull = ConvexHull(verts)
    tris = hull.simplices.copy()
    # outward winding (hull_mass sums signed tetrahedra): flip the facets whose vertex order disagrees with the facet normal
    n = np.cross(verts[tris[:, 1]] - verts[tris[:, 0]], verts[tris[:, 2]] - verts[tris[:, 0]])
    flip = np.einsum("ij,ij->i", n, hull.equations[:, :3]) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    return ShapeRec(SHAPE_CONVEX, pose7() if pose is None else pose, np.zeros(3), vertices=verts, triangles=tris, **kw)


def rotate_inertia(I, R):
    return R @ I @ R.T


def combine_mass(parts):
    """parts: list of (m, com(3), I(3x3 about com)) in one frame -> combined."""
    m = sum(p[0] for p in parts)
    if m <= 0:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    com = sum(p[0] * np.asarray(p[1]) for p in parts) / m
    I = np.zeros((3, 3))
    for pm, pc, pI in parts:
        d = np.asarray(pc) - com
        I += pI + pm * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    return m, com, I


def sym6(I):
    return [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]


# ----------------------------------------------------------------------------- records
@dataclass
class ShapeRec:
    type: int
    pose: np.ndarray
    size: np.ndarray = field(default_factory=lambda: np.zeros(3))
    vertices: Optional[np.ndarray] = None
    triangles: Optional[np.ndarray] = None
    mu: float = 0.3
    patch_radius: float = 0.0
    density: float = DEFAULT_DENSITY
    groups: Sequence[int] = (1, 1, 0, 0)
    per_env_size: Optional[np.ndarray] = None  # [n_envs,3]
    per_env_pose: Optional[np.ndarray] = None  # [n_envs,7]
    color: Sequence[float] = (0.7, 0.7, 0.7, 1.0)
    collide: bool = True
    visual: bool = True

    def mass_props(self, size=None):
        size = self.size if size is None else size
        if self.type == SHAPE_BOX:
            m, c, I = box_mass(size, self.density)
        elif self.type == SHAPE_SPHERE:
            m, c, I = sphere_mass(size[0], self.density)
        elif self.type == SHAPE_CAPSULE:
            m, c, I = capsule_mass(size[0], size[1], self.density)
        elif self.type == SHAPE_CONVEX:
            m, c, I = hull_mass(self.vertices, self.triangles, self.density)
        else:
            return 0.0, np.zeros(3), np.zeros((3, 3))
        R = qmat(self.pose[3:])
        return m, self.pose[:3] + R @ c, rotate_inertia(I, R)

    def bound(self, size=None, pose=None):
        size = self.size if size is None else size
        pose = self.pose if pose is None else pose
        if self.type == SHAPE_BOX:
            return np.array([*pose[:3], float(np.linalg.norm(size))])
        if self.type == SHAPE_SPHERE:
            return np.array([*pose[:3], float(size[0])])
        if self.type == SHAPE_CAPSULE:
            return np.array([*pose[:3], float(size[0] + size[1])])
        if self.type == SHAPE_CONVEX:
            v = np.asarray(self.vertices)
            c = v.mean(0)
            r = float(np.linalg.norm(v - c, axis=1).max())
            return np.array([*(pose[:3] + qrot(pose[3:], c)), r])
        return np.array([0, 0, 0, 1e9])


@dataclass
class ActorRec:
    name: str
    body_type: str  # dynamic | kinematic | static
    shapes: List[ShapeRec]
    initial_pose: np.ndarray
    mass: Optional[float] = None
    com: Optional[np.ndarray] = None
    inertia: Optional[np.ndarray] = None
    linear_damping: float = 0.0
    angular_damping: float = 0.05  # PhysX default angular damping
    disable_gravity: bool = False
    hidden: bool = False


@dataclass
class LinkInfo:
    name: str
    row: int
    dof: int  # abody (global dof index) or -(art+1)
    offset: np.ndarray


class ArticulationRec:
    def __init__(self, name, robot: dict, root_pose, link_mu=None, disable_gravity=True, drive=None):
        self.name = name
        self.robot = robot
        self.root_pose = np.asarray(root_pose, dtype=np.float64)
        self.link_mu = link_mu or {}
        self.link_patch: Dict[str, float] = {}
        self.disable_gravity = disable_gravity
        self.drive: Dict[str, Sequence[float]] = drive or {}  # joint name -> (kp, kd, force_limit)
        self.joint_friction: Dict[str, float] = {}
        self.link_groups: Dict[str, Sequence[int]] = {}
        self.extra_disabled: List[Sequence[int]] = []


def load_robot(name: str) -> dict:
    with open(os.path.join(ASSET_DIR, "robots", name + ".json")) as f:
        return json.load(f)


# ----------------------------------------------------------------------------- ctypes mirror of B2SModel
_I = C.POINTER(C.c_int32)
_U = C.POINTER(C.c_uint32)
_F = C.POINTER(C.c_float)


class B2SModelStruct(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("n_art", C.c_int32), ("n_dof", C.c_int32), ("n_link", C.c_int32), ("n_fb", C.c_int32),
        ("n_shape", C.c_int32), ("n_pair", C.c_int32), ("n_hull", C.c_int32), ("n_hull_verts", C.c_int32), ("n_eq", C.c_int32),
        ("n_ov_shape", C.c_int32), ("n_ov_fb", C.c_int32), ("max_contacts", C.c_int32), ("max_manifolds", C.c_int32), ("n_pos_iters", C.c_int32),
        ("n_vel_iters", C.c_int32), ("max_dof_per_art", C.c_int32),
        ("dt", C.c_float), ("gravity", C.c_float * 3), ("contact_offset", C.c_float), ("rest_offset", C.c_float),
        ("max_depen_vel", C.c_float), ("contact_hertz", C.c_float), ("contact_zeta", C.c_float), ("margin_min", C.c_float),
        ("dof_parent", _I), ("dof_art", _I), ("dof_type", _I), ("dof_T0", _F), ("dof_axis", _F), ("dof_mass", _F),
        ("dof_com", _F), ("dof_inertia", _F), ("dof_gravity", _F), ("dof_limit", _F), ("dof_drive", _F), ("dof_passive", _F),
        ("dof_anc_mask", _U),
        ("link_dof", _I), ("link_offset", _F),
        ("art_root_pose", _F), ("art_dof_start", _I), ("art_link_start", _I),
        ("eq_dof", _I), ("eq_param", _F),
        ("fb_type", _I), ("fb_mass", _F), ("fb_com", _F), ("fb_inertia", _F), ("fb_damping", _F), ("fb_gravity", _F),
        ("fb_init_pose", _F), ("fb_ov", _I),
        ("shape_type", _I), ("shape_owner_kind", _I), ("shape_owner", _I), ("shape_row", _I), ("shape_pose", _F),
        ("shape_size", _F), ("shape_hull", _I), ("shape_mu", _F), ("shape_bound", _F), ("shape_ov", _I), ("shape_patch", _F),
        ("hull_offset", _I), ("hull_verts", _F), ("hull_aabb", _F),
        ("pair_a", _I), ("pair_b", _I),
        ("ov_shape_size", _F), ("ov_shape_pose", _F), ("ov_shape_bound", _F), ("ov_fb_mass", _F),
    ]


@dataclass
class SimParams:
    """Mirror of SimConfig/SceneConfig defaults (mani_skill/utils/structs/types.py:38-97)."""
    sim_freq: int = 100
    control_freq: int = 20
    gravity: Sequence[float] = (0.0, 0.0, -9.81)
    contact_offset: float = 0.02
    rest_offset: float = 0.0
    solver_position_iterations: int = 15
    solver_velocity_iterations: int = 1
    static_friction: float = 0.3
    max_contacts: int = 64
    max_manifolds: int = 24
    max_depenetration_velocity: float = 3.0
    contact_hertz: float = 30.0
    contact_zeta: float = 10.0
    margin_min: float = 0.005
    max_joint_velocity: float = 100.0   # PxArticulationJointReducedCoordinate maxJointVelocity default (rad/s | m/s); 0 = unlimited

    @classmethod
    def from_config(cls, cfg=None) -> "SimParams":
        """Accepts what the reference's tasks pass as `sim_config` / `_default_sim_config` (mani_skill/utils/structs/types.py:78-97): a
        flat dict of this class's fields, or the nested `SimConfig` layout -- `scene_config` (SceneConfig: gravity, contact_offset,
        rest_offset, solver_position_iterations, solver_velocity_iterations, ...), `default_materials_config` (static_friction) and
        `gpu_memory_config` (PhysX buffer capacities, which have no counterpart here: the per-sub-scene capacities are `max_contacts` /
        `max_manifolds`) -- as dicts or dataclass instances.  Unknown keys are ignored like dacite's non-strict mode would; a key that is
        known but not honoured by this backend (enable_ccd, enable_tgs=False, ...) is kept out silently as well."""
        import dataclasses
        if cfg is None:
            return cls()
        if isinstance(cfg, cls):
            return cfg
        if dataclasses.is_dataclass(cfg):
            cfg = dataclasses.asdict(cfg)
        flat = {}
        names = {f.name for f in dataclasses.fields(cls)}
        for k, v in dict(cfg).items():
            if dataclasses.is_dataclass(v):
                v = dataclasses.asdict(v)
            if k in ("scene_config", "default_materials_config", "gpu_memory_config") and isinstance(v, dict):
                for kk, vv in v.items():
                    if kk in names:
                        flat[kk] = vv
            elif k in names:
                flat[k] = v
        return cls(**flat)


class CompiledModel:
    """Flat tables + name maps. ``struct()`` gives a ctypes B2SModel whose pointers alias arrays kept alive here."""

    def __init__(self):
        self.arrays: Dict[str, np.ndarray] = {}
        self.scalars: Dict[str, object] = {}
        self.link_rows: Dict[str, Dict[str, int]] = {}
        self.link_info: Dict[str, List[LinkInfo]] = {}
        self.actor_rows: Dict[str, int] = {}
        self.actor_fb: Dict[str, int] = {}
        self.art_index: Dict[str, int] = {}
        self.dof_names: Dict[str, List[str]] = {}
        self.art_dof_start: List[int] = []
        self.visuals: List[dict] = []
        self.hull_tris: List[np.ndarray] = []
        self.actor_seg_id: Dict[str, int] = {}
        self.link_seg_id: Dict[str, Dict[str, int]] = {}

    def struct(self) -> B2SModelStruct:
        s = B2SModelStruct()
        for k, v in self.scalars.items():
            if k == "gravity":
                s.gravity = (C.c_float * 3)(*v)
            else:
                setattr(s, k, v)
        for k, v in self.arrays.items():
            ftype = dict(B2SModelStruct._fields_)[k]
            if v.size == 0:
                v = np.zeros(1, dtype=v.dtype)
                self.arrays[k] = v
            setattr(s, k, v.ctypes.data_as(ftype))
        return s

    @property
    def n_rows(self):
        return self.scalars["n_link"] + self.scalars["n_fb"]


def _collide_groups(ga, gb):
    return (((ga[0] & gb[1]) | (ga[1] & gb[0])) != 0) and ((ga[2] & gb[2]) == 0)


class SceneDesc:
    """One sub-scene prototype (+ optional per-env shape parameters)."""

    def __init__(self, n_envs: int, sim: Optional[SimParams] = None):
        self.n_envs = n_envs
        self.sim = sim or SimParams()
        self.articulations: List[ArticulationRec] = []
        self.actors: List[ActorRec] = []

    def add_articulation(self, rec: ArticulationRec):
        self.articulations.append(rec)
        return rec

    def add_actor(self, rec: ActorRec):
        self.actors.append(rec)
        return rec

    # ------------------------------------------------------------------ compile
    def compile(self) -> CompiledModel:
        sim = self.sim
        cm = CompiledModel()
        N = self.n_envs
        dof_parent, dof_art, dof_type, dof_T0, dof_axis = [], [], [], [], []
        dof_mass_parts: List[list] = []
        dof_gravity, dof_limit, dof_drive, dof_passive = [], [], [], []
        link_dof, link_offset = [], []
        art_root, art_dof_start, art_link_start = [], [0], [0]
        eq_dof, eq_param = [], []
        shapes = []  # dict(rec, owner_kind, owner, row, art, link_idx, groups, size, pose)
        seg_next = 1
        row = 0
        for ai, art in enumerate(self.articulations):
            cm.art_index[art.name] = ai
            links = art.robot["links"]
            link_abody = [None] * len(links)  # (dof index or -(ai+1), offset pose7 in abody frame)
            names = []
            cm.link_rows[art.name] = {}
            cm.link_info[art.name] = []
            cm.link_seg_id[art.name] = {}
            jname_to_dof = {}
            for li, L in enumerate(links):
                J = L["joint"]
                jpose = pose7(J["p"], J["q"])
                if L["parent"] < 0:
                    ab, off = -(ai + 1), IDENTITY.copy()
                else:
                    pab, poff = link_abody[L["parent"]]
                    if J["type"] == "fixed":
                        ab, off = pab, pose_mul(poff, jpose)
                    else:
                        d = len(dof_parent)
                        dof_parent.append(pab if pab >= 0 else -1)
                        dof_art.append(ai)
                        dof_type.append(JOINT_PRISMATIC if J["type"] == "prismatic" else JOINT_REVOLUTE)
                        dof_T0.append(pose_mul(poff, jpose))
                        ax = np.asarray(J["axis"], dtype=np.float64)
                        dof_axis.append(ax / np.linalg.norm(ax))
                        dof_mass_parts.append([])
                        dof_gravity.append(0.0 if art.disable_gravity else 1.0)
                        lo, hi = J["lower"], J["upper"]
                        if J["type"] == "revolute_unwrapped":
                            lo, hi = -1e30, 1e30
                        dof_limit.append([lo, hi])
                        kp, kd, fl = art.drive.get(J["name"], (0.0, 0.0, 1e10))
                        dof_drive.append([kp, kd, fl, float(self.sim.max_joint_velocity)])
                        dof_passive.append([0.0, art.joint_friction.get(J["name"], 0.0), 0.0, 0.0])
                        names.append(J["name"])
                        jname_to_dof[J["name"]] = d
                        # "frame_offset": the link frame inside the frame the joint moves (joints anchored off the link origin,
                        # building.ArticulationBuilder); baked URDF robots have the two coincide
                        ab, off = d, (np.asarray(L["frame_offset"], dtype=np.float64) if "frame_offset" in L else IDENTITY.copy())
                link_abody[li] = (ab, off)
                link_dof.append(ab)
                link_offset.append(off)
                cm.link_rows[art.name][L["name"]] = row
                cm.link_info[art.name].append(LinkInfo(L["name"], row, ab, off))
                cm.link_seg_id[art.name][L["name"]] = seg_next
                # mass into abody
                if ab >= 0 and L["mass"] > 0:
                    R = qmat(off[3:])
                    Il = np.array([[L["inertia"][0], L["inertia"][3], L["inertia"][4]],
                                   [L["inertia"][3], L["inertia"][1], L["inertia"][5]],
                                   [L["inertia"][4], L["inertia"][5], L["inertia"][2]]])
                    dof_mass_parts[ab].append((L["mass"], off[:3] + R @ np.asarray(L["com"]), rotate_inertia(Il, R)))
                groups = art.link_groups.get(L["name"], (1, 1, 0, 0))
                for s in L["collisions"]:
                    if "rec" in s:   # a ready ShapeRec in the link frame (maniskill_b200/compat/compile.py): collision and / or visual
                        rec = ShapeRec(**{**s["rec"].__dict__})
                        rec.pose = pose_mul(off, s["rec"].pose)
                        if rec.per_env_pose is not None:
                            rec.per_env_pose = np.stack([pose_mul(off, p) for p in rec.per_env_pose])
                        shapes.append(dict(rec=rec, owner_kind=OWNER_LINK, owner=ab, row=row, art=ai, link=li, seg=seg_next, hidden=False))
                        continue
                    lp = pose_mul(off, pose7(s["p"], s["q"]))
                    mu = art.link_mu.get(L["name"], sim.static_friction)
                    if s["type"] == "box":
                        rec = ShapeRec(SHAPE_BOX, lp, np.asarray(s["half_size"], dtype=np.float64), mu=mu)
                    elif s["type"] == "sphere":
                        rec = ShapeRec(SHAPE_SPHERE, lp, np.array([s["radius"], 0, 0]), mu=mu)
                    elif s["type"] == "convex":
                        rec = ShapeRec(SHAPE_CONVEX, lp, np.zeros(3), vertices=np.asarray(s["vertices"]),
                                       triangles=np.asarray(s["triangles"]), mu=mu)
                    else:
                        continue
                    rec.groups = groups
                    rec.patch_radius = art.link_patch.get(L["name"], 0.0)
                    rec.color = (0.85, 0.85, 0.88, 1.0)
                    shapes.append(dict(rec=rec, owner_kind=OWNER_LINK, owner=ab, row=row, art=ai, link=li, seg=seg_next,
                                       hidden=False))
                seg_next += 1
                row += 1
            # fixed tendons from <mimic> (articulation_builder.py:161-200: stiffness 1e5)
            for L in links:
                J = L["joint"]
                if "mimic" in J and J["name"] in jname_to_dof and J["mimic"]["joint"] in jname_to_dof:
                    eq_dof.append([jname_to_dof[J["mimic"]["joint"]], jname_to_dof[J["name"]]])
                    eq_param.append([J["mimic"]["multiplier"], J["mimic"]["offset"], 1e5, 0.0])
            art_root.append(art.root_pose)
            art_dof_start.append(len(dof_parent))
            art_link_start.append(len(link_dof))
            cm.dof_names[art.name] = names
        n_dof = len(dof_parent)
        assert n_dof <= 32, "n_dof per env is limited to 32"
        n_link = len(link_dof)
        # merged abody mass props
        dof_mass, dof_com, dof_inertia = [], [], []
        for parts in dof_mass_parts:
            m, c, I = combine_mass(parts)
            dof_mass.append(m)
            dof_com.append(c)
            dof_inertia.append(sym6(I))
        anc = []
        for i in range(n_dof):
            mask = 1 << i
            p = dof_parent[i]
            while p >= 0:
                mask |= 1 << p
                p = dof_parent[p]
            anc.append(mask)
        # ---- actors
        fb_type, fb_mass, fb_com, fb_inertia, fb_damping, fb_gravity, fb_init, fb_ov = [], [], [], [], [], [], [], []
        ov_fb = []
        for act in self.actors:
            seg = seg_next
            seg_next += 1
            cm.actor_seg_id[act.name] = seg
            if act.body_type == "static":
                for s in act.shapes:
                    rec = ShapeRec(**{**s.__dict__})
                    rec.pose = pose_mul(act.initial_pose, s.pose)
                    shapes.append(dict(rec=rec, owner_kind=OWNER_STATIC, owner=0, row=-1, art=-1, link=-1, seg=seg, hidden=act.hidden))
                cm.actor_rows[act.name] = -1
                continue
            b = len(fb_type)
            cm.actor_fb[act.name] = b
            cm.actor_rows[act.name] = n_link + b
            fb_type.append(BODY_DYNAMIC if act.body_type == "dynamic" else BODY_KINEMATIC)
            per_env = any(s.per_env_size is not None for s in act.shapes) and act.body_type == "dynamic"
            if act.mass is not None:
                m, c, I = act.mass, np.zeros(3) if act.com is None else act.com, act.inertia
            else:
                m, c, I = combine_mass([s.mass_props() for s in act.shapes if s.collide])
            if m <= 0:
                m, I = 1.0, np.eye(3)
            fb_mass.append(m)
            fb_com.append(c)
            fb_inertia.append(sym6(I))
            fb_damping.append([act.linear_damping, act.angular_damping])
            fb_gravity.append(0.0 if act.disable_gravity else 1.0)
            fb_init.append(act.initial_pose)
            if per_env:
                tab = np.zeros((N, 10))
                for e in range(N):
                    parts = []
                    for s in act.shapes:
                        if not s.collide:
                            continue
                        sz = s.per_env_size[e] if s.per_env_size is not None else s.size
                        parts.append(s.mass_props(sz))
                    me, ce, Ie = combine_mass(parts)
                    tab[e, 0] = me
                    tab[e, 1:4] = ce
                    tab[e, 4:] = sym6(Ie)
                fb_ov.append(len(ov_fb))
                ov_fb.append(tab)
            else:
                fb_ov.append(-1)
            for s in act.shapes:
                shapes.append(dict(rec=s, owner_kind=OWNER_BODY, owner=b, row=n_link + b, art=-1, link=-1, seg=seg, hidden=act.hidden))
        n_fb = len(fb_type)
        # ---- shapes -> tables (collision shapes only); visuals recorded for the renderer
        hull_offset, hull_verts, hull_aabb = [0], [], []
        st, sok, so, srow, spose, ssize, shull, smu, sbound, sov, spatch = [], [], [], [], [], [], [], [], [], [], []
        ov_size, ov_pose, ov_bound = [], [], []
        col_meta = []
        for sh in shapes:
            rec: ShapeRec = sh["rec"]
            hull_id = -1
            if rec.type == SHAPE_CONVEX:
                hull_id = len(hull_offset) - 1
                hull_verts.extend(np.asarray(rec.vertices, dtype=np.float32).tolist())
                hull_offset.append(len(hull_verts))
                hv_ = np.asarray(rec.vertices, dtype=np.float64)
                hull_aabb.append(np.concatenate([(hv_.max(0) + hv_.min(0)) / 2, (hv_.max(0) - hv_.min(0)) / 2]))
                cm.hull_tris.append(np.asarray(rec.triangles, dtype=np.int32))
            if rec.visual:
                cm.visuals.append(dict(type=rec.type, row=sh["row"], pose=rec.pose, size=rec.size, hull=hull_id,
                                       color=rec.color, seg=sh["seg"], hidden=sh["hidden"],
                                       per_env_size=rec.per_env_size, per_env_pose=rec.per_env_pose))
            if not rec.collide:
                continue
            st.append(rec.type)
            sok.append(sh["owner_kind"])
            so.append(sh["owner"])
            srow.append(sh["row"])
            spose.append(rec.pose)
            ssize.append(rec.size)
            shull.append(hull_id)
            smu.append(rec.mu)
            spatch.append(rec.patch_radius)
            sbound.append(rec.bound())
            if rec.per_env_size is not None or rec.per_env_pose is not None:
                sov.append(len(ov_size))
                sz = rec.per_env_size if rec.per_env_size is not None else np.tile(rec.size, (N, 1))
                ps = rec.per_env_pose if rec.per_env_pose is not None else np.tile(rec.pose, (N, 1))
                ov_size.append(np.asarray(sz, dtype=np.float64))
                ov_pose.append(np.asarray(ps, dtype=np.float64))
                ov_bound.append(np.stack([rec.bound(sz[e], ps[e]) for e in range(N)]))
            else:
                sov.append(-1)
            col_meta.append(sh)
        n_shape = len(st)
        # ---- candidate pairs
        disabled = set()
        for ai, art in enumerate(self.articulations):
            for a, b in list(art.robot.get("disabled_collision_pairs", [])) + list(art.extra_disabled):
                disabled.add((ai, min(a, b), max(a, b)))
        pair_a, pair_b = [], []
        pairs = []

        def immovable(sh):
            if sh["owner_kind"] == OWNER_STATIC:
                return True
            if sh["owner_kind"] == OWNER_LINK:
                return sh["owner"] < 0
            return fb_type[sh["owner"]] == BODY_KINEMATIC

        for i in range(n_shape):
            for j in range(i + 1, n_shape):
                A, B = col_meta[i], col_meta[j]
                if immovable(A) and immovable(B):
                    continue
                if A["owner_kind"] == B["owner_kind"] and A["owner"] == B["owner"] and A["owner_kind"] != OWNER_STATIC:
                    continue
                if not _collide_groups(A["rec"].groups, B["rec"].groups):
                    continue
                if A["art"] >= 0 and A["art"] == B["art"]:
                    la, lb = A["link"], B["link"]
                    if (A["art"], min(la, lb), max(la, lb)) in disabled:
                        continue
                    # parent/child abodies never collide (PhysX articulation default)
                    oa, ob = A["owner"], B["owner"]
                    pa = dof_parent[oa] if oa >= 0 else None
                    pb = dof_parent[ob] if ob >= 0 else None
                    root_a, root_b = oa < 0, ob < 0
                    if (not root_a and not root_b and (pa == ob or pb == oa)) or (root_a and not root_b and pb == -1) or (
                            root_b and not root_a and pa == -1):
                        continue
                prio = 0 if (A["owner_kind"] == OWNER_BODY or B["owner_kind"] == OWNER_BODY) else 1
                pairs.append((prio, i, j))
        pairs.sort()
        for _, i, j in pairs:
            pair_a.append(i)
            pair_b.append(j)
        f32 = lambda x, shape=None: np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(shape) if shape else np.asarray(x, dtype=np.float32))
        i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
        n_ov_shape, n_ov_fb = len(ov_size), len(ov_fb)
        A = cm.arrays
        A["dof_parent"] = i32(dof_parent); A["dof_art"] = i32(dof_art); A["dof_type"] = i32(dof_type)
        A["dof_T0"] = f32(dof_T0).reshape(-1); A["dof_axis"] = f32(dof_axis).reshape(-1); A["dof_mass"] = f32(dof_mass)
        A["dof_com"] = f32(dof_com).reshape(-1); A["dof_inertia"] = f32(dof_inertia).reshape(-1); A["dof_gravity"] = f32(dof_gravity)
        A["dof_limit"] = f32(dof_limit).reshape(-1); A["dof_drive"] = f32(dof_drive).reshape(-1); A["dof_passive"] = f32(dof_passive).reshape(-1)
        A["dof_anc_mask"] = np.ascontiguousarray(np.asarray(anc, dtype=np.uint32))
        A["link_dof"] = i32(link_dof); A["link_offset"] = f32(link_offset).reshape(-1)
        A["art_root_pose"] = f32(art_root).reshape(-1); A["art_dof_start"] = i32(art_dof_start); A["art_link_start"] = i32(art_link_start)
        A["eq_dof"] = i32(eq_dof).reshape(-1); A["eq_param"] = f32(eq_param).reshape(-1)
        A["fb_type"] = i32(fb_type); A["fb_mass"] = f32(fb_mass); A["fb_com"] = f32(fb_com).reshape(-1)
        A["fb_inertia"] = f32(fb_inertia).reshape(-1); A["fb_damping"] = f32(fb_damping).reshape(-1); A["fb_gravity"] = f32(fb_gravity)
        A["fb_init_pose"] = f32(fb_init).reshape(-1); A["fb_ov"] = i32(fb_ov)
        A["shape_type"] = i32(st); A["shape_owner_kind"] = i32(sok); A["shape_owner"] = i32(so); A["shape_row"] = i32(srow)
        A["shape_pose"] = f32(spose).reshape(-1); A["shape_size"] = f32(ssize).reshape(-1); A["shape_hull"] = i32(shull)
        A["shape_mu"] = f32(smu); A["shape_bound"] = f32(sbound).reshape(-1); A["shape_ov"] = i32(sov); A["shape_patch"] = f32(spatch)
        A["hull_offset"] = i32(hull_offset); A["hull_verts"] = f32(hull_verts).reshape(-1); A["hull_aabb"] = f32(hull_aabb).reshape(-1)
        A["pair_a"] = i32(pair_a); A["pair_b"] = i32(pair_b)
        A["ov_shape_size"] = f32(np.stack(ov_size, 1) if ov_size else np.zeros(0)).reshape(-1)
        A["ov_shape_pose"] = f32(np.stack(ov_pose, 1) if ov_pose else np.zeros(0)).reshape(-1)
        A["ov_shape_bound"] = f32(np.stack(ov_bound, 1) if ov_bound else np.zeros(0)).reshape(-1)
        A["ov_fb_mass"] = f32(np.stack(ov_fb, 1) if ov_fb else np.zeros(0)).reshape(-1)
        max_dof = max([art_dof_start[i + 1] - art_dof_start[i] for i in range(len(self.articulations))] + [0])
        cm.scalars.update(
            n_envs=N, n_art=len(self.articulations), n_dof=n_dof, n_link=n_link, n_fb=n_fb, n_shape=n_shape, n_pair=len(pair_a),
            n_hull=len(hull_offset) - 1, n_hull_verts=len(hull_verts), n_eq=len(eq_dof), n_ov_shape=n_ov_shape, n_ov_fb=n_ov_fb,
            max_contacts=sim.max_contacts, max_manifolds=sim.max_manifolds, n_pos_iters=sim.solver_position_iterations, n_vel_iters=sim.solver_velocity_iterations,
            max_dof_per_art=max_dof, dt=1.0 / sim.sim_freq, gravity=tuple(sim.gravity), contact_offset=sim.contact_offset,
            rest_offset=sim.rest_offset, max_depen_vel=sim.max_depenetration_velocity, contact_hertz=sim.contact_hertz, contact_zeta=sim.contact_zeta, margin_min=sim.margin_min,
        )
        cm.art_dof_start = art_dof_start
        return cm
