"""`sapien.physx.PhysxGpuSystem`-shaped facade over one b200sim world (the GPU-only entry points).

The reference drives the simulator through a small set of `px.*` calls and `px.cuda_*` buffers
(mani_skill/envs/scene.py:902-986 `_gpu_apply_all` / `_gpu_fetch_all`, :741-801 contact queries, :379-380 `px.step()`,
mani_skill/utils/structs/{actor,link,articulation}.py index them with `gpu_pose_index` / `gpu_index`).  This module gives a
world built through the C-ABI (maniskill_b200/backend.py `World`) those names, argument meanings and buffer semantics, so that code
written against `px` -- the reference's `ManiSkillScene` in particular -- finds the calls it makes:

    px = PhysxGpuSystem(world)
    px.cuda_rigid_body_data.torch()[body.gpu_pose_index, :7] = pose;  px.gpu_apply_rigid_dynamic_data()
    px.cuda_articulation_target_qpos.torch()[art.gpu_index, :9] = t;  px.gpu_apply_articulation_target_position()
    px.step();  px.gpu_fetch_rigid_dynamic_data();  px.gpu_fetch_articulation_qpos() ...
    q = px.gpu_create_contact_pair_impulse_query([(a, b), ...]);  px.gpu_query_contact_pair_impulses(q);  q.cuda_impulses.torch()

Tensors returned by `.torch()` are persistent aliases of the world's device buffers (the caller mutates them in place, then calls
`gpu_apply_*`; ownership stays with the world).  What the builders do in the reference (entities, components, `gpu_init`) has
happened when the world was compiled (maniskill_b200/model.py); the scene-building half of the `sapien` module is not part of this
facade.  No computation happens here: every method is one C-ABI call on torch's current stream.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .backend import BUF_LINK, BUF_QACC, BUF_QF, BUF_QPOS, BUF_QVEL, BUF_RIGID, BUF_ROOT_POSE, BUF_TARGET_QPOS, BUF_TARGET_QVEL

ANY_BODY = -2  # include/b200sim.h B2S_ANY_BODY


class CudaArray:
    """`px.cuda_*` object: `.torch()` hands out the persistent, aliasing tensor (sapien's CudaArray; scene.py:902-948)."""

    def __init__(self, tensor: torch.Tensor):
        self._t = tensor
        self.shape = tuple(tensor.shape)

    def torch(self) -> torch.Tensor:
        return self._t


class BodyHandle:
    """What the reference reads off a `PhysxRigidBodyComponent` on this path: `gpu_pose_index` = its row in `cuda_rigid_body_data`
    (actor.py:352-354, link.py:251-269), `gpu_index` (the same row here), the owning sub-scene and the prototype row."""

    __slots__ = ("env", "row", "gpu_pose_index", "gpu_index", "name")

    def __init__(self, env: int, row: int, n_rows: int, name: str):
        self.env, self.row, self.name = env, row, name
        self.gpu_pose_index = self.gpu_index = env * n_rows + row

    def __repr__(self):
        return f"BodyHandle({self.name!r}, env={self.env}, gpu_pose_index={self.gpu_pose_index})"


class ArticulationHandle:
    """`PhysxArticulation.gpu_index` = its row in the `cuda_articulation_*` buffers (articulation.py:873-896)."""

    __slots__ = ("env", "index", "gpu_index", "name", "dof")

    def __init__(self, env: int, index: int, n_art: int, name: str, dof: int):
        self.env, self.index, self.name, self.dof = env, index, name, dof
        self.gpu_index = env * n_art + index


class ContactImpulseQuery:
    """Result object of `gpu_create_contact_{pair,body}_impulse_query`: `.cuda_impulses.torch()` is `[n, 3]`, refreshed by the
    matching `gpu_query_*` call (scene.py:771-781, base.py:116-136)."""

    def __init__(self, key, env_index: torch.Tensor, col_index: torch.Tensor, out: torch.Tensor):
        self._key, self._env, self._col = key, env_index, col_index
        self.cuda_impulses = CudaArray(out)


class PhysxGpuSystem:
    def __init__(self, world, body_names: Sequence[str] = (), articulation_names: Sequence[Tuple[str, int]] = ()):
        """world: maniskill_b200.backend.World (or an object with its interface).  body_names: name per prototype row of
        `cuda_rigid_body_data`; articulation_names: (name, dof) per prototype articulation -- both optional, for the handles."""
        self._w = world
        N, R, A = world.n_envs, world.n_rows, max(world.n_art, 1)
        self.device = world.device
        self.cuda_rigid_body_data = CudaArray(world.rigid_body_data)                 # [N * rows, 13] pos3 quat_wxyz4 linvel3 angvel3
        self.cuda_articulation_qpos = CudaArray(world.qpos)                          # [N * arts, max_dof]
        self.cuda_articulation_qvel = CudaArray(world.qvel)
        self.cuda_articulation_qacc = CudaArray(world.qacc)
        self.cuda_articulation_qf = CudaArray(world.qf)
        self.cuda_articulation_target_qpos = CudaArray(world.target_qpos)
        self.cuda_articulation_target_qvel = CudaArray(world.target_qvel)
        names = list(body_names) or [f"body{r}" for r in range(R)]
        if len(names) != R:
            raise ValueError(f"{len(names)} body names for {R} rows per sub-scene")
        self.bodies: List[List[BodyHandle]] = [[BodyHandle(e, r, R, names[r]) for r in range(R)] for e in range(N)]
        arts = list(articulation_names) or [(f"articulation{i}", world.max_dof) for i in range(world.n_art)]
        self.articulations: List[List[ArticulationHandle]] = [[ArticulationHandle(e, i, A, n, d) for i, (n, d) in enumerate(arts)] for e in range(N)]
        self._dt = float(world.cm.scalars["dt"])

    # ---- PhysxSystem.timestep (sapien_env.py:1227): fixed when the world was compiled
    @property
    def timestep(self) -> float:
        return self._dt

    @timestep.setter
    def timestep(self, dt: float):
        if abs(float(dt) - self._dt) > 1e-12:
            raise RuntimeError(f"the world was compiled with timestep {self._dt}; rebuild it to change the timestep (asked for {dt})")

    def gpu_init(self):
        """Buffers exist from world creation on; kept so that `px.gpu_init()` (scene.py:902) is a valid call."""

    # ---- scene.py:379-380
    def step(self):
        self._w.step(1, 0)

    # ---- scene.py:950-966 (and sapien_env.py:1118-1121 for the drive targets)
    def gpu_apply_rigid_dynamic_data(self):
        self._w.apply(BUF_RIGID)

    def gpu_apply_articulation_qpos(self):
        self._w.apply(BUF_QPOS)

    def gpu_apply_articulation_qvel(self):
        self._w.apply(BUF_QVEL)

    def gpu_apply_articulation_qf(self):
        self._w.apply(BUF_QF)

    def gpu_apply_articulation_root_pose(self):
        """Root pose travels through the root link's row of `cuda_rigid_body_data` (articulation.py:821-859)."""
        self._w.apply(BUF_ROOT_POSE)

    def gpu_apply_articulation_root_velocity(self):
        """Fixed-base articulations only (`fix_root_link=True`, base_agent.py:71,174): there is no root velocity to apply."""

    def gpu_apply_articulation_target_position(self):
        self._w.apply(BUF_TARGET_QPOS)

    def gpu_apply_articulation_target_velocity(self):
        self._w.apply(BUF_TARGET_QVEL)

    def gpu_update_articulation_kinematics(self):
        self._w.update_kinematics()

    # ---- scene.py:968-986
    def gpu_fetch_rigid_dynamic_data(self):
        self._w.fetch(BUF_RIGID)

    def gpu_fetch_articulation_link_pose(self):
        self._w.fetch(BUF_LINK)

    def gpu_fetch_articulation_link_velocity(self):
        """Link poses and velocities share one row of `cuda_rigid_body_data`; `gpu_fetch_articulation_link_pose` wrote both."""

    def gpu_fetch_articulation_qpos(self):
        self._w.fetch(BUF_QPOS)

    def gpu_fetch_articulation_qvel(self):
        self._w.fetch(BUF_QVEL)

    def gpu_fetch_articulation_qacc(self):
        self._w.fetch(BUF_QACC)

    def gpu_fetch_articulation_target_qpos(self):
        self._w.fetch(BUF_TARGET_QPOS)

    def gpu_fetch_articulation_target_qvel(self):
        self._w.fetch(BUF_TARGET_QVEL)

    # ---- scene.py:741-801, base.py:116-136
    def _make_query(self, pairs: Sequence[Tuple[BodyHandle, object]]) -> ContactImpulseQuery:
        uniq, env_idx, col_idx = {}, [], []
        for a, b in pairs:
            rb = ANY_BODY if b is ANY_BODY else (-1 if b is None else b.row)
            if rb >= 0 and b.env != a.env:
                raise RuntimeError(f"contact query between bodies of different sub-scenes: {a!r}, {b!r}")
            env_idx.append(a.env)
            col_idx.append(uniq.setdefault((a.row, rb), len(uniq)))
        key = self._w.create_contact_query(list(uniq.keys()))
        dev = self.device
        return ContactImpulseQuery(key, torch.tensor(env_idx, dtype=torch.int64, device=dev), torch.tensor(col_idx, dtype=torch.int64, device=dev),
                                   torch.zeros((len(pairs), 3), dtype=torch.float32, device=dev))

    def _run_query(self, q: ContactImpulseQuery):
        out = self._w.query_contact_impulses(q._key)        # [n_envs, n_unique_row_pairs, 3]
        q.cuda_impulses.torch().copy_(out[q._env, q._col])

    def gpu_create_contact_pair_impulse_query(self, body_pairs: Sequence[Tuple[BodyHandle, BodyHandle]]) -> ContactImpulseQuery:
        """Sum of the solver's contact impulses (world frame, acting on the first body) between the two bodies of each pair;
        `None` as the second body stands for the static geometry."""
        return self._make_query(list(body_pairs))

    def gpu_query_contact_pair_impulses(self, query: ContactImpulseQuery):
        self._run_query(query)

    def gpu_create_contact_body_impulse_query(self, bodies: Sequence[BodyHandle]) -> ContactImpulseQuery:
        """Net contact impulse on each body (over everything touching it)."""
        return self._make_query([(b, ANY_BODY) for b in bodies])

    def gpu_query_contact_body_impulses(self, query: ContactImpulseQuery):
        self._run_query(query)
