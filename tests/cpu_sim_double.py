"""TEST DOUBLE for `sapien.physx.PhysxCpuSystem`: the reference's `sim_backend="physx_cpu"` path over the CPU oracle.

The product's `PhysxCpuSystem()` raises (there is no CPU simulation in b200sim).  SURVEY.md section 8(c) asks for the reference's own semantic tests
re-run "against the oracle through unchanged ManiSkill code", and BASELINE.json's configs[0] is PickCube-v1 with num_envs=1 on the CPU backend: with this
double installed the reference's CPU code path (per-object getters / setters, `px.step()`, `px.get_contacts()`) runs with the ORACLE where SAPIEN's
CPU PhysX would be, so that the reference's CPU-vs-GPU tests compare the oracle with the device code through the reference's own Python.

How: sapien's CPU objects are live (a getter returns the simulated value).  Here the shim's recording objects stay the source of truth between calls;
`step()` pushes all of them into an `OracleBackendWorld` (compiled from the same records as the batched world), steps it and pulls everything back; a link
pose / velocity getter pushes, runs forward kinematics and pulls.  Only classes of the shim are patched, and only while `installed()` is active.
"""
import contextlib

import numpy as np

from oracle_world import OracleBackendWorld


def _p7(pose):
    return np.concatenate([np.asarray(pose.p, dtype=np.float64), np.asarray(pose.q, dtype=np.float64)])


@contextlib.contextmanager
def installed(precision="f32"):
    import sapien
    from sapien import physx

    import maniskill_b200.compat as compat
    from maniskill_b200.compat import compile as _c

    class _NoBody:
        entity = None

    class OracleCpuSystem(physx.PhysxSystem):
        """Bookkeeping of PhysxGpuSystem (component lists per sub-scene), one sub-scene, compiled lazily at the first use that needs the simulation."""

        def __init__(self, *a, **kw):
            super().__init__()
            self.device = sapien.Device("cpu")
            self._offsets = {}
            self.rigid_dynamic_components, self.rigid_static_components, self.articulation_link_components = [], [], []
            self._by_scene = {}
            self._sim = None          # OracleBackendWorld
            self._compiled = None

        _unregister_component = physx.PhysxGpuSystem._unregister_component
        get_rigid_dynamic_components = lambda self: self.rigid_dynamic_components
        get_rigid_static_components = lambda self: self.rigid_static_components
        get_articulation_link_components = lambda self: self.articulation_link_components

        def _register_component(self, component, scene):
            if self._sim is not None:
                raise RuntimeError("the CPU test double compiles its world at the first simulation call; entities cannot be added afterwards")
            physx.PhysxGpuSystem._register_component(self, component, scene)

        # ---- objects <-> oracle world
        def _ensure(self):
            if self._sim is None:
                saved = compat.WORLD_FACTORY
                compat.WORLD_FACTORY = lambda cm, dev: OracleBackendWorld(cm, precision)
                try:
                    self._compiled = _c.compile_system(self, physx._CONFIG)
                finally:
                    compat.WORLD_FACTORY = saved
                self._sim = self._compiled.world
            return self._sim

        def _components(self):
            return [c for comps in self._by_scene.values() for c in comps]

        def _push(self):
            w = self._ensure()
            body = w.body_view()
            arts = {}
            for c in self._components():
                if c._row < 0:
                    continue
                if isinstance(c, physx.PhysxArticulationLinkComponent):
                    arts[id(c.articulation)] = c.articulation
                    if c.parent is None:
                        body[0, c._row, :7] = _t(_p7(c.articulation._root_pose))
                    continue
                body[0, c._row, :7] = _t(_p7(c._pose))
                body[0, c._row, 7:10] = _t(c.__dict__.get("linear_velocity", np.zeros(3)))
                body[0, c._row, 10:13] = _t(c.__dict__.get("angular_velocity", np.zeros(3)))
            for art in arts.values():
                a, n = art.gpu_index, art.dof
                w.qpos[a, :n] = _t(art.qpos)
                w.qvel[a, :n] = _t(art.qvel)
                w.qf[a, :n] = _t(art.qf)
                tq = np.concatenate([np.asarray(j.drive_target, dtype=np.float64).reshape(-1) for j in art.active_joints] or [np.zeros(0)])
                tv = np.concatenate([np.asarray(j.drive_velocity_target, dtype=np.float64).reshape(-1) for j in art.active_joints] or [np.zeros(0)])
                w.target_qpos[a, :n] = _t(tq)
                w.target_qvel[a, :n] = _t(tv)
            w.apply()

        def _pull(self):
            w = self._sim
            body = w.body_view().numpy()
            arts = {}
            for c in self._components():
                if c._row < 0:
                    continue
                r = body[0, c._row]
                p = sapien.Pose(r[:3].copy(), r[3:7].copy())
                c._pose = p
                if c.entity is not None:
                    c.entity._pose = p
                c.__dict__["linear_velocity"] = r[7:10].astype(np.float32)
                c.__dict__["angular_velocity"] = r[10:13].astype(np.float32)
                if isinstance(c, physx.PhysxArticulationLinkComponent):
                    arts[id(c.articulation)] = c.articulation
            for art in arts.values():
                a, n = art.gpu_index, art.dof
                art._qpos = w.qpos[a, :n].numpy().copy()
                art._qvel = w.qvel[a, :n].numpy().copy()
                art._qacc = w.qacc[a, :n].numpy().copy()

        def _refresh_links(self):
            """forward kinematics of the objects' current (root pose, qpos, qvel)"""
            self._push()
            self._sim.update_kinematics()
            self._pull()

        # ---- the CPU system's API
        def step(self):
            self._push()
            self._sim.step(1)
            self._pull()

        def get_contacts(self):
            if self._sim is None:
                return []
            by_row = {c._row: c for c in self._components() if c._row >= 0}
            out = {}
            for row in self._sim.contacts(0):
                key = (int(row[0]), int(row[1]))
                pt = physx.PhysxContactPoint()
                pt.position, pt.normal, pt.separation, pt.impulse = row[2:5].copy(), row[5:8].copy(), float(row[8]), row[9:12].copy()
                if key not in out:
                    c = out[key] = physx.PhysxContact()
                    c.bodies = [by_row.get(key[0], _NoBody), by_row.get(key[1], _NoBody)]
                    c.shapes, c.points = [None, None], []
                out[key].points.append(pt)
            return list(out.values())

        def set_scene_offset(self, scene, offset):
            self._offsets[id(scene)] = np.asarray(offset, dtype=np.float32).reshape(3)

        def get_scene_offset(self, scene):
            return self._offsets.get(id(scene), np.zeros(3, dtype=np.float32))

    def _t(x):
        import torch
        return torch.as_tensor(np.asarray(x, dtype=np.float32))

    def _live(component):
        s = getattr(component, "_system", None)
        return s if isinstance(s, OracleCpuSystem) else None

    Link, Art = physx.PhysxArticulationLinkComponent, physx.PhysxArticulation
    _MISSING = object()
    patched = [(physx, "PhysxCpuSystem"), (Art, "qvel"), (Art, "qf"), (Art, "qacc"), (Art, "get_qvel"), (Art, "set_qvel"), (Art, "get_qf"), (Art, "set_qf"), (Art, "get_qacc"),
               (Art, "compute_passive_force"), (Link, "_body_pose"), (Link, "pose"), (Link, "linear_velocity"), (Link, "angular_velocity"),
               (physx.PhysxRigidBaseComponent, "entity_pose")]
    saved = {(cls, k): cls.__dict__.get(k, _MISSING) for cls, k in patched}

    # ---- articulation: generalized velocities / forces are stored (the GPU objects return zeros: live values are in the cuda buffers)
    def _stored(name):
        def get(self):
            v = getattr(self, "_" + name, None)
            return np.zeros(self.dof, dtype=np.float32) if v is None else v

        def put(self, v):
            setattr(self, "_" + name, np.asarray(v, dtype=np.float32).reshape(-1).copy())
        return property(get, put)

    Art.qvel, Art.qf, Art.qacc = _stored("qvel"), _stored("qf"), _stored("qacc")
    Art.get_qvel, Art.get_qf, Art.get_qacc = (lambda self: self.qvel), (lambda self: self.qf), (lambda self: self.qacc)
    Art.set_qvel = lambda self, v: setattr(self, "qvel", v)
    Art.set_qf = lambda self, v: setattr(self, "qf", v)
    Art.compute_passive_force = lambda self, gravity=True, coriolis_and_centrifugal=True: np.zeros(self.dof, dtype=np.float32)

    # ---- links: pose / velocity follow (root pose, qpos, qvel)
    link_pose_setter = Link.__dict__["pose"].fset

    def link_body_pose(self):
        s = _live(self)
        if s is not None:
            s._refresh_links()
        return self._pose

    def link_velocity(name):
        def get(self):
            s = _live(self)
            if s is not None:
                s._refresh_links()
            return self.__dict__.get(name, np.zeros(3, dtype=np.float32))

        def put(self, v):
            self.__dict__[name] = np.asarray(v, dtype=np.float32).reshape(3)
        return property(get, put)

    Link._body_pose = link_body_pose
    Link.pose = property(lambda self: link_body_pose(self), link_pose_setter)
    Link.linear_velocity, Link.angular_velocity = link_velocity("linear_velocity"), link_velocity("angular_velocity")
    physx.PhysxRigidBaseComponent.entity_pose = property(lambda self: self.entity.pose if self.entity is not None else self.pose)
    physx.PhysxCpuSystem = OracleCpuSystem
    try:
        yield OracleCpuSystem
    finally:
        for (cls, k), v in saved.items():
            if v is _MISSING:
                if k in cls.__dict__:
                    delattr(cls, k)
            else:
                setattr(cls, k, v)
