"""`PhysxGpuSystem.gpu_init()`: the recorded sapien scene graphs -> ONE batched b200sim world.

The reference builds N python object graphs, one `sapien.Scene` per sub-scene (mani_skill/envs/sapien_env.py:1186-1210,
mani_skill/utils/building/actor_builder.py:234-245, articulation_builder.py:139-211).  Here sub-scene 0 is the PROTOTYPE: its
articulations and actors become the `ArticulationRec` / `ActorRec` tables of maniskill_b200/model.py, compiled once and instantiated
`n_envs` times on the device.  The other sub-scenes must have the same structure (same entities, component kinds and shape types in the
same order -- what ManiSkill's builders produce); where they differ in shape size / shape pose (per-sub-scene geometry,
mani_skill/envs/tasks/tabletop/peg_insertion_side.py:114-191) the differences become the per-env override tables, and their initial
poses are written into the world after it is created.  Every component then learns its row: `gpu_pose_index = env * rows + row`
(mani_skill/utils/structs/actor.py:352-354, link.py:251-269), `gpu_index` of an articulation = `env * n_art + a`
(articulation.py:873-896).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from ..model import (SHAPE_BOX, SHAPE_CAPSULE, SHAPE_CONVEX, SHAPE_PLANE, SHAPE_SPHERE, ActorRec, ArticulationRec, SceneDesc, ShapeRec, SimParams,
                     combine_mass, cylinder_shape, pose7, pose_inv, pose_mul, qmat, qrot, rotate_inertia)

_CYL_X_TO_PLANE = np.array([0.7071068, 0, -0.7071068, 0])   # rotates +x (plane normal of SHAPE_PLANE) onto +z


def _p7(pose) -> np.ndarray:
    return pose.raw() if pose is not None else pose7()


def _strip(name: str) -> str:
    """'scene-3_cube' -> 'cube'; 'scene-3-panda_panda_link0' -> 'panda_link0' is handled by the caller (it knows the articulation name)."""
    if name.startswith("scene-"):
        head, sep, tail = name.partition("_")
        if sep:
            return tail
    return name


# ------------------------------------------------------------------------------------------------ shapes
def collision_shape_rec(shape, visual: bool) -> Optional[ShapeRec]:
    """sapien collision shape -> ShapeRec in the frame of its body."""
    from sapien import physx
    lp = _p7(shape.local_pose)
    m = shape.physical_material
    kw = dict(mu=float(m.dynamic_friction), patch_radius=max(float(shape.patch_radius), float(shape.min_patch_radius)), density=float(shape.density),
              groups=tuple(int(g) for g in shape.collision_groups), collide=True, visual=visual)
    k = shape.kind
    if k == "plane":
        return ShapeRec(SHAPE_PLANE, lp, **kw)
    if k == "box":
        return ShapeRec(SHAPE_BOX, lp, np.asarray(shape.half_size, dtype=np.float64), **kw)
    if k == "sphere":
        return ShapeRec(SHAPE_SPHERE, lp, np.array([shape.radius, 0.0, 0.0]), **kw)
    if k == "capsule":
        return ShapeRec(SHAPE_CAPSULE, lp, np.array([shape.radius, shape.half_length, 0.0]), **kw)
    if k == "cylinder":
        return cylinder_shape(shape.radius, shape.half_length, lp, **kw)
    if k == "convex_mesh":
        return ShapeRec(SHAPE_CONVEX, lp, np.zeros(3), vertices=np.asarray(shape.vertices, dtype=np.float64), triangles=np.asarray(shape.triangles), **kw)
    raise NotImplementedError(f"collision shape kind {k}")


def render_shape_recs(shape) -> List[ShapeRec]:
    """sapien render shape -> visual-only ShapeRecs.  Triangle meshes are drawn as the convex hull of each part (<= 64 vertices, the
    part's base colour); flat meshes (floors) as a half-space."""
    from .. import meshio
    lp = _p7(shape.local_pose)
    kw = dict(collide=False, visual=True)
    color = tuple(shape.material.effective_color()) if shape.material is not None else (0.7, 0.7, 0.7, 1.0)
    k = shape.kind
    if k == "box":
        return [ShapeRec(SHAPE_BOX, lp, np.asarray(shape.half_size, dtype=np.float64), color=color, **kw)]
    if k == "sphere":
        return [ShapeRec(SHAPE_SPHERE, lp, np.array([shape.radius, 0.0, 0.0]), color=color, **kw)]
    if k in ("capsule", "cylinder"):
        return [cylinder_shape(shape.radius, shape.half_length + (shape.radius if k == "capsule" else 0.0), lp, color=color, **kw)]
    if k == "plane":
        return [ShapeRec(SHAPE_PLANE, lp, color=color, **kw)]
    if k == "mesh":
        out = []
        for part in shape.parts:
            v = np.asarray(part.vertices, dtype=np.float64)
            if len(v) < 3:
                continue
            pc = tuple(part.material.effective_color()) if part.material is not None else color
            ext = v.max(0) - v.min(0)
            flat = int(np.argmin(ext))
            if ext[flat] < 1e-6 * max(1.0, float(ext.max())):
                # a flat sheet (the ground's tiled quad mesh, ground.py:62-120): a half-space through it, normal along the flat axis
                # pointing to the side its triangles face
                tri = v[np.asarray(part.triangles)[0]]
                n = np.cross(tri[1] - tri[0], tri[2] - tri[0])
                sign = 1.0 if n[flat] >= 0 else -1.0
                axis = np.zeros(3)
                axis[flat] = sign
                from sapien.math import shortest_rotation
                q = shortest_rotation([1, 0, 0], axis).astype(np.float64)
                c = v.mean(0)
                out.append(ShapeRec(SHAPE_PLANE, pose_mul(lp, pose7(c, q)), color=pc, **kw))
                continue
            try:
                hv, ht = meshio.cook_hull(v)
            except Exception:
                continue
            out.append(ShapeRec(SHAPE_CONVEX, lp, np.zeros(3), vertices=hv, triangles=np.asarray(ht), color=pc, **kw))
        return out
    return []


def shape_world_points(shape, body_pose) -> np.ndarray:
    """Corner / vertex cloud of a collision shape in the world frame (AABB queries)."""
    rec = collision_shape_rec(shape, False)
    return _rec_points(rec, _p7(body_pose))


def render_shape_world_points(shape, body_pose) -> np.ndarray:
    pts = [_rec_points(r, _p7(body_pose)) for r in render_shape_recs(shape)]
    return np.concatenate(pts) if pts else np.zeros((1, 3))


def _rec_points(rec: ShapeRec, body7) -> np.ndarray:
    T = pose_mul(body7, rec.pose)
    R = qmat(T[3:])
    if rec.type == SHAPE_BOX:
        c = np.array([[(1 if i & 1 else -1), (1 if i & 2 else -1), (1 if i & 4 else -1)] for i in range(8)]) * rec.size
    elif rec.type == SHAPE_SPHERE:
        c = np.array([[(1 if i & 1 else -1), (1 if i & 2 else -1), (1 if i & 4 else -1)] for i in range(8)]) * rec.size[0]
    elif rec.type == SHAPE_CAPSULE:
        c = np.array([[sx * (rec.size[0] + rec.size[1]), sy * rec.size[0], sz * rec.size[0]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    elif rec.type == SHAPE_CONVEX:
        c = np.asarray(rec.vertices)
    else:
        c = np.zeros((1, 3))
    return c @ R.T + T[:3]


def body_mass_props(body):
    """(mass, cmass_local_pose, principal inertia) of a rigid body from its collision shapes and their densities."""
    import sapien
    parts = [collision_shape_rec(s, False).mass_props() for s in body.collision_shapes if s.kind != "plane"]
    m, c, I = combine_mass(parts) if parts else (0.0, np.zeros(3), np.zeros((3, 3)))
    if m <= 0:
        return 1.0, sapien.Pose(), np.ones(3)
    w, V = np.linalg.eigh(I)
    if np.linalg.det(V) < 0:
        V[:, 2] *= -1
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = V, c
    return float(m), sapien.Pose(T), w


def _explicit_mass(body):
    """-> (mass, com(3), inertia 3x3 about the com in the body frame) or None when the body computes it from its shapes."""
    if body._mass is None:
        return None
    cp = _p7(body._cmass_local_pose) if body._cmass_local_pose is not None else pose7()
    R = qmat(cp[3:])
    I = np.asarray(body._inertia, dtype=np.float64) if body._inertia is not None else np.ones(3) * 1e-6
    return float(body._mass), cp[:3], rotate_inertia(np.diag(I), R)


# ------------------------------------------------------------------------------------------------ the compile step
class Compiled:
    def __init__(self):
        self.world = None
        self.facade = None
        self.cm = None
        self.system = None
        self.art_names: List[str] = []
        self.joint_dof: Dict[int, int] = {}   # id(prototype joint) -> global dof index

    def update_drive(self, joint, stiffness, damping, force_limit, mode):
        """Drive gains after gpu_init: the gains are one table shared by all sub-scenes.  Accepted when they equal what was compiled (the
        controllers re-apply their gains on every control-mode (re)set); different values need a reconfigure."""
        if abs(joint.stiffness - stiffness) < 1e-9 * max(1.0, abs(stiffness)) and abs(joint.damping - damping) < 1e-9 * max(1.0, abs(damping)) and \
                (joint.force_limit == force_limit or abs(joint.force_limit - force_limit) < 1e-6 * max(1.0, abs(force_limit))):
            return
        raise RuntimeError("joint drive gains are compiled into the batched world at gpu_init(); create the environment with the control mode "
                           "it will use (or reconfigure) instead of changing gains afterwards")

    def sync_poses_to_objects(self):
        """`px.sync_poses_gpu_to_cpu()`: entity poses of every sub-scene from the device buffers."""
        import sapien
        data = self.world.rigid_body_data.detach().cpu().numpy()
        for comps in self.system._by_scene.values():
            for c in comps:
                if c.gpu_pose_index >= 0:
                    p = sapien.Pose(data[c.gpu_pose_index, :3], data[c.gpu_pose_index, 3:7])
                    c._pose = p
                    if c.entity is not None:
                        c.entity._pose = p


def _scene_components(system, scene):
    return system._by_scene.get(id(scene), [])


def _split(components):
    """-> (articulations [[links...]...] in order of appearance, free bodies in order)."""
    from sapien import physx
    arts, art_links, bodies = [], {}, []
    for c in components:
        if isinstance(c, physx.PhysxArticulationLinkComponent):
            key = id(c.articulation)
            if key not in art_links:
                art_links[key] = []
                arts.append(art_links[key])
            art_links[key].append(c)
        else:
            bodies.append(c)
    for links in arts:
        links.sort(key=lambda l: l.index)
    return arts, bodies


def _render_body(component):
    from sapien import render
    e = component.entity
    return e.find_component_by_type(render.RenderBodyComponent) if e is not None else None


def _body_shapes(component, others: list, N: int) -> List[ShapeRec]:
    """ShapeRecs of one prototype body: its collision shapes (with per-env size / pose tables where the matching bodies of the other
    sub-scenes differ) + the visual shapes of its render body."""
    recs = []
    rb = _render_body(component)
    # plane collisions exist in the first sub-scene only (the reference's builder adds one plane per pose, actor_builder.py:75-86): the
    # prototype's plane serves every sub-scene; the remaining shapes are matched by position
    solid = lambda c: [x for x in c.collision_shapes if x.kind != "plane"]
    other_solids = [solid(o) for o in others]
    si = -1
    for s in component.collision_shapes:
        rec = collision_shape_rec(s, visual=False)
        if s.kind == "plane":
            recs.append(rec)
            continue
        si += 1
        if others:
            sizes, poses, differ = [rec.size], [rec.pose], False
            for o, osol in zip(others, other_solids):
                if si >= len(osol) or osol[si].kind != s.kind:
                    raise NotImplementedError(f"sub-scenes differ in structure (shapes of '{component.name}'); only sizes / poses may differ")
                r2 = collision_shape_rec(osol[si], False)
                if r2.type == SHAPE_CONVEX and (len(r2.vertices) != len(rec.vertices) or not np.allclose(r2.vertices, rec.vertices, atol=1e-7)):
                    raise NotImplementedError(f"per-sub-scene convex meshes are not supported ('{component.name}')")
                differ |= (not np.allclose(r2.size, rec.size, atol=1e-9)) or (not np.allclose(r2.pose, rec.pose, atol=1e-9))
                sizes.append(r2.size)
                poses.append(r2.pose)
            if differ:
                rec.per_env_size, rec.per_env_pose = np.stack(sizes), np.stack(poses)
        recs.append(rec)
    if rb is not None:
        for vi, rs in enumerate(rb.render_shapes):
            vrecs = render_shape_recs(rs)
            if others and len(vrecs) == 1 and vrecs[0].type in (SHAPE_BOX, SHAPE_SPHERE):
                sizes, poses, differ = [vrecs[0].size], [vrecs[0].pose], False
                for o in others:
                    orb = _render_body(o)
                    if orb is None or vi >= len(orb.render_shapes):
                        differ = False
                        break
                    r2 = render_shape_recs(orb.render_shapes[vi])
                    if len(r2) != 1 or r2[0].type != vrecs[0].type:
                        differ = False
                        break
                    differ |= (not np.allclose(r2[0].size, vrecs[0].size, atol=1e-9)) or (not np.allclose(r2[0].pose, vrecs[0].pose, atol=1e-9))
                    sizes.append(r2[0].size)
                    poses.append(r2[0].pose)
                if differ and len(sizes) == N:
                    vrecs[0].per_env_size, vrecs[0].per_env_pose = np.stack(sizes), np.stack(poses)
            recs.extend(vrecs)
    return recs


def compile_system(system, config) -> Compiled:
    import torch
    from sapien import physx
    from .. import compat
    from ..physx_shim import PhysxGpuSystem as Facade
    scenes = list(system.scenes)
    N = len(scenes)
    if N == 0:
        raise RuntimeError("gpu_init() without any sapien.Scene on the system")
    sim = SimParams(
        sim_freq=int(round(1.0 / system.timestep)), control_freq=int(round(1.0 / system.timestep)), gravity=tuple(float(g) for g in config["scene"]["gravity"]),
        contact_offset=float(config["shape"]["contact_offset"]), rest_offset=float(config["shape"]["rest_offset"]),
        solver_position_iterations=int(config["body"]["solver_position_iterations"]), solver_velocity_iterations=int(config["body"]["solver_velocity_iterations"]),
        static_friction=float(config["material"]["dynamic_friction"]))
    desc = SceneDesc(N, sim)
    per_scene = [_scene_components(system, s) for s in scenes]
    proto = per_scene[0]
    for e, comps in enumerate(per_scene[1:], 1):
        if len(comps) != len(proto) or any(type(a) is not type(b) for a, b in zip(comps, proto)):
            raise NotImplementedError(f"sub-scene {e} differs in structure from sub-scene 0 ({len(comps)} vs {len(proto)} bodies): the b200sim backend "
                                      "instantiates ONE prototype; only shape sizes / poses and initial poses may differ between sub-scenes")
    arts, bodies = _split(proto)
    index_of = {id(c): k for k, c in enumerate(proto)}
    out = Compiled()
    out.system = system
    # ---------------- articulations
    for links in arts:
        art = links[0].articulation
        aname = _strip(art.name) if art.name else f"articulation{len(out.art_names)}"
        out.art_names.append(aname)
        root = links[0]
        if root.joint.type not in ("fixed",):
            raise NotImplementedError(f"articulation '{aname}' has a free root link; the b200sim backend simulates fixed-base articulations "
                                      "(mobile bases are modelled with joints, like the reference's Fetch)")
        prefix = f"scene-0-{aname}_"
        lname = lambda l: l.name[len(prefix):] if l.name.startswith(prefix) else _strip(l.name)
        jname = lambda j: j.name[len(prefix):] if j.name.startswith(prefix) else _strip(j.name)
        link_index = {id(l): i for i, l in enumerate(links)}
        robot_links = []
        drive, jfric, link_groups, link_mu, link_patch = {}, {}, {}, {}, {}
        mimic_of = {}
        for t in art.tendons:
            chain, co = t["chain"], t["coefficients"]
            if len(chain) != 3 or abs(co[2]) < 1e-12:
                raise NotImplementedError("fixed tendons other than the two-joint mimic coupling are not supported")
            a, b = chain[1], chain[2]
            mimic_of[id(b)] = dict(joint=jname(a.joint), multiplier=-co[1] / co[2], offset=t["rest_length"] / co[2])
        for l in links:
            j = l.joint
            others = [per_scene[e][index_of[id(l)]] for e in range(1, N)]
            if j.type == "fixed" or l.parent is None:
                jt = "fixed"
                T = pose_mul(_p7(j.pose_in_parent), pose_inv(_p7(j.pose_in_child))) if l.parent is not None else pose7()
                jd = dict(name=jname(j), type="fixed", p=T[:3].tolist(), q=T[3:].tolist(), axis=[1, 0, 0], lower=0, upper=0, effort=0, damping=0, friction=0)
                frame_offset = None
            elif j.type in ("revolute", "revolute_unwrapped", "continuous", "prismatic"):
                jt = "prismatic" if j.type == "prismatic" else ("revolute" if j.type == "revolute" else "revolute_unwrapped")
                pip = _p7(j.pose_in_parent)
                lim = np.asarray(j.limit, dtype=np.float64).reshape(-1)
                lo, hi = (float(lim[0]), float(lim[1])) if lim.size == 2 else (-1e30, 1e30)
                lo = -1e30 if not np.isfinite(lo) else lo
                hi = 1e30 if not np.isfinite(hi) else hi
                jd = dict(name=jname(j), type=jt, p=pip[:3].tolist(), q=pip[3:].tolist(), axis=[1, 0, 0], lower=lo, upper=hi, effort=0, damping=0,
                          friction=float(j.friction))
                frame_offset = pose_inv(_p7(j.pose_in_child))
                fl = float(j.force_limit)
                drive[jname(j)] = (float(j.stiffness), float(j.damping), fl if np.isfinite(fl) and fl < 1e10 else 1e10)
                jfric[jname(j)] = float(j.friction)
                if id(l) in mimic_of:
                    jd["mimic"] = mimic_of[id(l)]
            else:
                raise NotImplementedError(f"joint type '{j.type}' ({j.name}) is not supported by the b200sim backend")
            em = _explicit_mass(l)
            if em is None:
                parts = [collision_shape_rec(s, False).mass_props() for s in l.collision_shapes if s.kind != "plane"]
                em = combine_mass(parts) if parts else (0.0, np.zeros(3), np.zeros((3, 3)))
            mass, com, I = em
            L = dict(name=lname(l), parent=-1 if l.parent is None else link_index[id(l.parent)], joint=jd, mass=float(mass), com=np.asarray(com).tolist(),
                     inertia=[I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]],
                     collisions=[dict(rec=r) for r in _body_shapes(l, others, N)])
            if frame_offset is not None:
                L["frame_offset"] = frame_offset.tolist()
            robot_links.append(L)
            if l.collision_shapes:
                s0 = l.collision_shapes[0]
                link_groups[L["name"]] = tuple(int(g) for g in s0.collision_groups)
        names = {L["name"]: i for i, L in enumerate(robot_links)}
        disabled = [[names[a], names[b]] for a, b in art.srdf_disabled_pairs if a in names and b in names]
        rec = ArticulationRec(aname, dict(name=aname, links=robot_links, disabled_collision_pairs=disabled), _p7(art.pose),
                              disable_gravity=all(l.disable_gravity for l in links), drive=drive)
        if not rec.disable_gravity and any(l.disable_gravity for l in links):
            raise NotImplementedError("gravity can be disabled for all links of an articulation or for none")
        rec.joint_friction = jfric
        rec.link_groups = link_groups
        desc.add_articulation(rec)
    # ---------------- actors
    actor_names = []
    for b in bodies:
        others = [per_scene[e][index_of[id(b)]] for e in range(1, N)]
        name = _strip(b.entity.name if b.entity is not None and b.entity.name else b.name) or f"actor{len(actor_names)}"
        k, base = 1, name
        while name in actor_names:
            name = f"{base}#{k}"
            k += 1
        actor_names.append(name)
        if isinstance(b, physx.PhysxRigidStaticComponent):
            body_type = "static"
        else:
            body_type = "kinematic" if b.kinematic else "dynamic"
            if all(b.locked_motion_axes):
                body_type = "kinematic"      # all six axes locked (base.py:343-354): the body only moves when its pose is set -- a kinematic body
            elif any(b.locked_motion_axes):
                raise NotImplementedError("partially locked motion axes are not supported by the b200sim backend (all six locked = kinematic is)")
        shapes = _body_shapes(b, others, N)
        pose0 = _p7(b.pose)
        if body_type == "static" and others:
            # static bodies have no row to carry a per-sub-scene pose: fold pose differences into the shapes
            poses = [pose0] + [_p7(o.pose) for o in others]
            if any(not np.allclose(p, pose0, atol=1e-9) for p in poses):
                for r in shapes:
                    loc = r.per_env_pose if r.per_env_pose is not None else np.tile(r.pose, (N, 1))
                    r.per_env_pose = np.stack([pose_mul(pose_mul(pose_inv(pose0), poses[e]), loc[e]) for e in range(N)])
                    if r.per_env_size is None:
                        r.per_env_size = np.tile(r.size, (N, 1))
        rec = ActorRec(name, body_type, shapes, pose0)
        if body_type != "static":
            em = _explicit_mass(b)
            if em is not None:
                rec.mass, rec.com, rec.inertia = em
            rec.linear_damping, rec.angular_damping, rec.disable_gravity = float(b.linear_damping), float(b.angular_damping), bool(b.disable_gravity)
        desc.add_actor(rec)
    cm = desc.compile()
    out.cm = cm
    dev_index = getattr(system.device, "cuda_id", 0)
    world = compat.make_world(cm, dev_index)
    out.world = world
    # ---------------- rows / indices of every component of every sub-scene
    n_rows, n_art = world.n_rows, max(world.n_art, 1)
    rows = {}
    for ai, links in enumerate(arts):
        aname = out.art_names[ai]
        prefix = f"scene-0-{aname}_"
        for l in links:
            ln = l.name[len(prefix):] if l.name.startswith(prefix) else _strip(l.name)
            rows[index_of[id(l)]] = (cm.link_rows[aname][ln], cm.link_seg_id[aname][ln])
    for b, name in zip(bodies, actor_names):
        rows[index_of[id(b)]] = (cm.actor_rows[name], cm.actor_seg_id[name])
    art_index = {}
    for ai, links in enumerate(arts):
        for l in links:
            art_index[index_of[id(l)]] = ai
    for e, comps in enumerate(per_scene):
        seen_art = set()
        for k, c in enumerate(comps):
            row, seg = rows[k]
            c._env, c._row = e, row
            c.env, c.row = e, row                     # what the contact-query facade reads (physx_shim.BodyHandle interface)
            c.gpu_pose_index = c.gpu_index = (e * n_rows + row) if row >= 0 else -1
            if c.entity is not None:
                c.entity.per_scene_id = seg           # segmentation id = per_scene_id (render/shaders.py:68-84)
            if k in art_index and id(c.articulation) not in seen_art:
                seen_art.add(id(c.articulation))
                c.articulation.gpu_index = e * n_art + art_index[k]
                c.articulation._env = e
    # ---------------- initial poses of the other sub-scenes (the prototype's are in the compiled model)
    body = world.body_view()
    dirty = False
    for e, comps in enumerate(per_scene):
        for k, c in enumerate(comps):
            row = rows[k][0]
            if row < 0:
                continue
            is_link = isinstance(c, physx.PhysxArticulationLinkComponent)
            if is_link and c.parent is not None:
                continue
            p = _p7(c.articulation.pose if is_link else c.pose)
            p0 = _p7(proto[k].articulation.pose if is_link else proto[k].pose)
            if e > 0 and not np.allclose(p, p0, atol=1e-9):
                body[e, row, :7] = torch.as_tensor(p, dtype=torch.float32, device=body.device)
                dirty = True
    if dirty:
        from ..backend import BUF_ALL, BUF_RIGID, BUF_ROOT_POSE
        world.apply(BUF_RIGID | BUF_ROOT_POSE)
        world.fetch(BUF_ALL)
    names = [""] * n_rows
    for k, c in enumerate(proto):
        if rows[k][0] >= 0:
            names[rows[k][0]] = c.name
    out.facade = Facade(world, names, [(n, len(cm.dof_names[n])) for n in out.art_names])
    return out


# ------------------------------------------------------------------------------------------------ cameras
def create_camera_group(render_group, cameras):
    """N `RenderCameraComponent`s (one per sub-scene, mani_skill/envs/scene.py:247-297) -> one camera of the batched rasteriser."""
    from sapien import physx
    from ..render import build_visual_table
    system = None
    for rs in render_group.systems:
        if rs.scene is not None and rs.scene.physx_system is not None:
            system = rs.scene.physx_system
            break
    if system is None or system._compiled is None:
        raise RuntimeError("create_camera_group before gpu_init()")
    comp = system._compiled
    cam = cameras[0]
    mount_row, local = -1, _p7(cam.local_pose)
    e = cam.entity
    body = e.find_component_by_type(physx.PhysxRigidBaseComponent) if e is not None else None
    if body is not None and body._row >= 0:
        mount_row = body._row
    elif e is not None:
        local = pose_mul(_p7(e.pose), local)
    for c in cameras[1:]:
        if (c.width, c.height) != (cam.width, cam.height) or abs(c.fx - cam.fx) > 1e-6 * cam.fx or abs(c.near - cam.near) > 1e-9:
            raise NotImplementedError("per-sub-scene camera intrinsics are not supported by the batched rasteriser")
    desc = dict(uid=cam.name, width=cam.width, height=cam.height, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, near=cam.near, far=cam.far,
                mount_row=mount_row, local_pose=[float(x) for x in local])
    if getattr(comp, "_visuals", None) is None:
        comp._visuals = build_visual_table(comp.cm, comp.world.n_envs)
    return comp.world.create_camera_group([desc], comp._visuals)
