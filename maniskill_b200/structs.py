"""Batched views over the backend's exposed buffers -- host-side mirror of mani_skill/utils/structs/.

Same names and semantics as the reference classes for the members the stock tasks use:
``Pose`` (pose.py:31-238), ``Actor`` (actor.py:341-403), ``Link`` (link.py:235-281), ``Articulation``
(articulation.py:723-921).  Setters honour ``scene._reset_mask`` exactly like the reference
(actor.py:389-391, articulation.py:784-787) so partial resets only touch the selected sub-scenes.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import utils as U


class Pose:
    """raw_pose [N,7] = p(3) + q(4, wxyz)."""

    def __init__(self, raw_pose: torch.Tensor):
        self.raw_pose = raw_pose

    @classmethod
    def create_from_pq(cls, p=None, q=None, device=None):
        if p is None:
            p = torch.zeros((1, 3), device=device)
        if q is None:
            q = torch.tensor([[1.0, 0, 0, 0]], device=device)
        p = U.to_tensor(p, device)
        q = U.to_tensor(q, device)
        if p.dim() == 1:
            p = p[None]
        if q.dim() == 1:
            q = q[None]
        n = max(p.shape[0], q.shape[0])
        return cls(torch.hstack([p.expand(n, 3), q.expand(n, 4)]).float())

    @classmethod
    def create(cls, pose, device=None):
        if isinstance(pose, Pose):
            return pose
        if hasattr(pose, "p") and hasattr(pose, "q") and not isinstance(pose, torch.Tensor):   # a single `sapien.Pose` (pose.py:128-135)
            pose = np.concatenate([np.asarray(pose.p, dtype=np.float32), np.asarray(pose.q, dtype=np.float32)])
        t = U.to_tensor(pose, device)
        if t.dim() == 1:
            t = t[None]
        return cls(t.float())

    @property
    def p(self):
        return self.raw_pose[..., :3]

    @property
    def q(self):
        return self.raw_pose[..., 3:]

    @property
    def shape(self):
        return self.raw_pose.shape

    @property
    def device(self):
        return self.raw_pose.device

    def __len__(self):
        return self.raw_pose.shape[0]

    def __getitem__(self, i):
        return Pose(self.raw_pose[i] if self.raw_pose[i].dim() == 2 else self.raw_pose[i][None])

    def __mul__(self, other: "Pose") -> "Pose":
        if not isinstance(other, Pose):   # a single `sapien.Pose` (building.Pose) on the right, as pose.py:187-199 accepts
            other = Pose.create(np.concatenate([np.asarray(other.p, dtype=np.float32), np.asarray(other.q, dtype=np.float32)]), self.raw_pose.device)
        a, b = self.raw_pose, other.raw_pose
        if b.shape[0] == 1 and a.shape[0] > 1:
            b = b.expand(a.shape[0], 7)
        elif a.shape[0] == 1 and b.shape[0] > 1:
            a = a.expand(b.shape[0], 7)
        p = a[..., :3] + U.quat_apply(a[..., 3:], b[..., :3])
        q = U.quat_mul(a[..., 3:], b[..., 3:])
        q = torch.where(q[..., :1] < 0, -q, q)  # quaternion_multiply standardises to a non-negative real part (pose.py:199)
        return Pose(torch.hstack([p, q]))

    def inv(self) -> "Pose":
        qi = U.quat_conj(self.q)
        return Pose(torch.hstack([-U.quat_apply(qi, self.p), qi]))

    def to_transformation_matrix(self) -> torch.Tensor:
        n = self.raw_pose.shape[0]
        top = torch.cat([U.quat_to_matrix(self.q), self.p.unsqueeze(-1)], dim=2)  # [n, 3, 4]
        key = (self.device, top.dtype)
        if key not in _LAST_ROW:
            _LAST_ROW[key] = torch.tensor([[[0.0, 0.0, 0.0, 1.0]]], device=self.device, dtype=top.dtype)
        return torch.cat([top, _LAST_ROW[key].expand(n, 1, 4)], dim=1)


_LAST_ROW = {}


def _sel(scene, idx):
    """idx[scene._reset_mask] without the boolean gather (and its device sync) when every sub-scene is selected."""
    return idx if getattr(scene, "_reset_all", False) else idx[scene._reset_mask]


class _BodyView:
    """Rows of rigid_body_data belonging to one named body across all sub-scenes."""

    def __init__(self, scene, name: str, row: int):
        self.scene = scene
        self.name = name
        self.row = row
        self._idx = torch.arange(scene.num_envs, device=scene.device) * scene.world.n_rows + row

    @property
    def _data(self):
        return self.scene.world.rigid_body_data

    @property
    def pose(self) -> Pose:
        return Pose(self._data[self._idx, :7])

    def get_pose(self):
        return self.pose

    @property
    def linear_velocity(self):
        return self._data[self._idx, 7:10]

    @property
    def angular_velocity(self):
        return self._data[self._idx, 10:13]

    def get_net_contact_impulses(self):
        """base.py:116-136: net contact impulse on this body over the last substep (sub-scene frame)."""
        return self.scene.get_net_contact_impulses(self)

    def get_net_contact_forces(self):
        return self.scene.get_net_contact_forces(self)

    def get_linear_velocity(self):
        return self.linear_velocity

    def get_angular_velocity(self):
        return self.angular_velocity


class Link(_BodyView):
    pass


class Actor(_BodyView):
    def __init__(self, scene, name, row, body_type, fb_index, initial_pose):
        super().__init__(scene, name, row)
        self.px_body_type = body_type
        self.fb_index = fb_index
        self.initial_pose = initial_pose
        self.hidden = False

    def is_static(self, lin_thresh=1e-2, ang_thresh=1e-1):
        """actor.py:220-227."""
        return torch.logical_and(torch.linalg.norm(self.linear_velocity, axis=1) <= lin_thresh, torch.linalg.norm(self.angular_velocity, axis=1) <= ang_thresh)

    def set_pose(self, pose):
        pose = Pose.create(pose, self.scene.device)
        rows = _sel(self.scene, self._idx)
        raw = pose.raw_pose
        if raw.shape[0] == 1:
            raw = raw.expand(rows.shape[0], 7)
        self._data[rows, :7] = raw
        self.scene._dirty |= self.scene.BUF_RIGID

    def set_linear_velocity(self, v):
        v = U.to_tensor(v, self.scene.device)
        self._data[_sel(self.scene, self._idx), 7:10] = v
        self.scene._dirty |= self.scene.BUF_RIGID

    def set_angular_velocity(self, v):
        v = U.to_tensor(v, self.scene.device)
        self._data[_sel(self.scene, self._idx), 10:13] = v
        self.scene._dirty |= self.scene.BUF_RIGID

    def get_state(self):
        return self._data[self._idx, :13].clone()

    def set_state(self, state, env_idx=None):
        m = self.scene._reset_mask if env_idx is None else env_idx
        self._data[self._idx[m], :13] = state
        self.scene._dirty |= self.scene.BUF_RIGID


class Articulation:
    def __init__(self, scene, name: str, art_index: int, link_rows: Dict[str, int], dof_names: List[str], qlimits: np.ndarray):
        self.scene = scene
        self.name = name
        self.art_index = art_index
        self.links_map = {n: Link(scene, n, r) for n, r in link_rows.items()}
        self.links = list(self.links_map.values())
        self.root = self.links[0]
        self.dof_names = dof_names
        self.dof = len(dof_names)
        self.max_dof = scene.world.max_dof
        w = scene.world
        self._rows = torch.arange(scene.num_envs, device=scene.device) * max(w.n_art, 1) + art_index
        self.qlimits = torch.tensor(qlimits, dtype=torch.float32, device=scene.device)[None].expand(scene.num_envs, -1, -1)

    def get_links(self):
        return self.links

    # --- joint state (articulation.py:723-815)
    @property
    def qpos(self):
        return self.scene.world.qpos[self._rows, :self.dof]

    @property
    def qvel(self):
        return self.scene.world.qvel[self._rows, :self.dof]

    @property
    def qacc(self):
        return self.scene.world.qacc[self._rows, :self.dof]

    def get_qpos(self):
        return self.qpos

    def get_qvel(self):
        return self.qvel

    def get_qlimits(self):
        return self.qlimits

    def _masked_rows(self):
        return _sel(self.scene, self._rows)

    def set_qpos(self, q):
        q = U.to_tensor(q, self.scene.device)
        self.scene.world.qpos[self._masked_rows(), :self.dof] = q
        self.scene._dirty |= self.scene.BUF_QPOS

    def set_qvel(self, q):
        q = U.to_tensor(q, self.scene.device)
        self.scene.world.qvel[self._masked_rows(), :self.dof] = q
        self.scene._dirty |= self.scene.BUF_QVEL

    def set_qf(self, q):
        q = U.to_tensor(q, self.scene.device)
        self.scene.world.qf[self._masked_rows(), :self.dof] = q
        self.scene._dirty |= self.scene.BUF_QF

    def set_joint_drive_targets(self, targets, joint_indices):
        """articulation.py:873-896: write px.cuda_articulation_target_qpos[gx[mask], gy[mask]]."""
        rows = _sel(self.scene, self._rows)
        self.scene.world.target_qpos[rows[:, None], joint_indices[None, :]] = targets
        self.scene._dirty |= self.scene.BUF_TARGET_QPOS

    def set_joint_drive_velocity_targets(self, targets, joint_indices):
        rows = _sel(self.scene, self._rows)
        self.scene.world.target_qvel[rows[:, None], joint_indices[None, :]] = targets
        self.scene._dirty |= self.scene.BUF_TARGET_QVEL

    @property
    def pose(self):
        return self.root.pose

    def set_pose(self, pose):
        pose = Pose.create(pose, self.scene.device)
        rows = _sel(self.scene, self.root._idx)
        self.scene.world.rigid_body_data[rows, :7] = pose.raw_pose
        self.scene._dirty |= self.scene.BUF_ROOT_POSE

    def get_state(self):
        """13 + 2*dof per env (articulation.py:283-313): root pose/vel, qpos, qvel."""
        root = self.scene.world.rigid_body_data[self.root._idx, :13]
        return torch.hstack([root, self.qpos, self.qvel])

    def set_state(self, state, env_idx=None):
        m = self.scene._reset_mask if env_idx is None else env_idx
        self.scene.world.rigid_body_data[self.root._idx[m], :7] = state[:, :7]
        self.scene.world.qpos[self._rows[m], :self.dof] = state[:, 13:13 + self.dof]
        self.scene.world.qvel[self._rows[m], :self.dof] = state[:, 13 + self.dof:13 + 2 * self.dof]
        self.scene._dirty |= self.scene.BUF_ROOT_POSE | self.scene.BUF_QPOS | self.scene.BUF_QVEL
