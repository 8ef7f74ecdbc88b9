"""Batched forward kinematics, geometric Jacobian and the one-step damped least-squares IK of the end-effector controllers.

Host-side mirror of mani_skill/agents/controllers/utils/kinematics.py (`Kinematics`, GPU branch :197-275): the reference builds a
`pytorch_kinematics` serial chain from the URDF, takes `chain.jacobian(q)` (6 x n: linear rows first, then angular, expressed in
the root frame) and solves  (J^T J + 1e-4 I) dq = J^T delta_pose  (Levenberg-Marquardt, one iteration) or `pinv(J) delta_pose`.
Here the chain comes from the baked robot description (tools/bake_assets.py) and everything is plain batched torch.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch

from . import utils as U


class SerialChain:
    """Root link -> `end_link` of a baked robot (maniskill_b200/assets/robots/*.json): fixed and 1-dof joints, wxyz quaternions."""

    def __init__(self, robot: dict, end_link: str, device, dtype=torch.float32):
        names = [l["name"] for l in robot["links"]]
        if end_link not in names:
            raise KeyError(end_link)
        chain = []
        i = names.index(end_link)
        while i >= 0:
            chain.append(i)
            i = robot["links"][i]["parent"]
        chain.reverse()
        self.link_names = [names[i] for i in chain]
        self.joint_names: List[str] = []   # the moving joints along the chain, root first
        self._origin_p, self._origin_q, self._axis, self._kind = [], [], [], []
        for i in chain[1:]:  # the root link has no joint
            j = robot["links"][i]["joint"]
            self._origin_p.append(j["p"])
            self._origin_q.append(j["q"])
            self._axis.append(j["axis"])
            kind = {"fixed": 0, "revolute": 1, "continuous": 1, "prismatic": 2}[j["type"]]
            self._kind.append(kind)
            if kind:
                self.joint_names.append(j["name"])
        t = lambda a: torch.tensor(a, dtype=dtype, device=device)
        self.origin_p = t(self._origin_p)                       # [n_elem, 3]
        self.origin_R = U.quat_to_matrix(t(self._origin_q))      # [n_elem, 3, 3]
        self.axis = t(self._axis)                                # [n_elem, 3]
        moving = [i for i, k in enumerate(self._kind) if k]
        ax = self.axis[moving] if moving else torch.zeros((0, 3), dtype=dtype, device=device)
        K = torch.zeros((len(moving), 3, 3), dtype=dtype, device=device)   # cross-product matrices of the joint axes
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 2], ax[:, 1], ax[:, 2], -ax[:, 0], -ax[:, 1], ax[:, 0]
        self._K, self._K2 = K, K @ K
        self._revolute = torch.tensor([self._kind[i] == 1 for i in moving], dtype=torch.bool, device=device)
        self._eye = torch.eye(3, dtype=dtype, device=device)
        self.n_joints = len(self.joint_names)
        self.device, self.dtype = device, dtype

    def forward(self, q: torch.Tensor):
        """q [B, n_joints] -> end-link position [B,3], rotation matrix [B,3,3] and the Jacobian [B, 6, n_joints] in the root frame
        (rows 0-2: linear velocity of the end-link origin, rows 3-5: angular velocity), pytorch_kinematics convention.
        Rotations of all revolute joints are built at once (Rodrigues: I + sin K + (1 - cos) K^2); the walk down the chain is then
        a handful of small batched matrix products per joint."""
        B = q.shape[0]
        s, c = torch.sin(q), torch.cos(q)
        Rj = self._eye + s[..., None, None] * self._K + (1.0 - c)[..., None, None] * self._K2   # [B, n_joints, 3, 3]
        p = torch.zeros((B, 3), dtype=self.dtype, device=self.device)
        R = self._eye.expand(B, 3, 3)
        joint_p, joint_axis = [], []
        k = 0
        for i, kind in enumerate(self._kind):
            p = p + R @ self.origin_p[i]          # joint frame = parent link frame * origin
            R = R @ self.origin_R[i]
            if kind == 0:
                continue
            axis_w = R @ self.axis[i]
            joint_p.append(p)
            joint_axis.append(axis_w)
            if kind == 1:
                R = R @ Rj[:, k]
            else:
                p = p + axis_w * q[:, k:k + 1]
            k += 1
        if not joint_p:
            return p, R, torch.zeros((B, 6, 0), dtype=self.dtype, device=self.device)
        JP, JA = torch.stack(joint_p, dim=2), torch.stack(joint_axis, dim=2)        # [B, 3, n]
        lin_rev = torch.linalg.cross(JA, p.unsqueeze(2) - JP, dim=1)
        rev = self._revolute
        J = torch.cat([torch.where(rev, lin_rev, JA), torch.where(rev, JA, torch.zeros_like(JA))], dim=1)
        return p, R, J


class Kinematics:
    """kinematics.py:25-275 restricted to the GPU branch: IK of `end_link` over the controlled joints `joint_names` (the other
    moving joints of the chain are held, kinematics.py:170-187 `qmask`)."""

    def __init__(self, robot: dict, end_link: str, dof_names: Sequence[str], joint_names: Sequence[str], device, articulation=None):
        self.chain = SerialChain(robot, end_link, device)
        dof_names = list(dof_names)
        # chain joints as indices into the articulation's qpos; which of them are controlled
        self.chain_dof_idx = torch.tensor([dof_names.index(n) for n in self.chain.joint_names], dtype=torch.int64, device=device)
        ctrl = [n for n in self.chain.joint_names if n in set(joint_names)]
        if ctrl != list(joint_names):
            raise ValueError(f"controlled joints {list(joint_names)} must be the chain joints {self.chain.joint_names} in chain order")
        self.qmask = torch.tensor([n in set(joint_names) for n in self.chain.joint_names], dtype=torch.bool, device=device)
        self.device = device
        # on the CUDA world the Levenberg-Marquardt step is one kernel (include/b200sim.h b2s_ik_step); the torch code below stays the
        # reference it is tested against (tests/test_gpu_round2.py) and serves worlds without the entry point (host emulation)
        self._ik = None
        world = getattr(getattr(articulation, "scene", None), "world", None)
        if world is not None and hasattr(world, "create_ik") and os.environ.get("B2S_IK_KERNEL", "1") not in ("", "0"):
            c = self.chain
            moving = [k != 0 for k in c._kind]
            names = iter(c.joint_names)
            columns, controlled = [], []
            for mv in moving:
                n = next(names) if mv else None
                columns.append(dof_names.index(n) if mv else 0)
                controlled.append(1 if mv and n in set(joint_names) else 0)
            origin7 = [list(p) + list(q) for p, q in zip(c._origin_p, c._origin_q)]
            self._ik_world, self._ik_art = world, articulation
            self._ik_args = (origin7, c._axis, c._kind, columns, controlled)
            self._ik_handles = {}

    def fk(self, qpos: torch.Tensor):
        """End-link position and wxyz quaternion in the root frame."""
        p, R, _ = self.chain.forward(qpos[:, self.chain_dof_idx])
        return p, U.matrix_to_quat(R)

    def compute_ik(self, delta_pose: torch.Tensor, q0: torch.Tensor, solver_config: dict):
        """delta_pose [B,6] = (translation, XYZ Euler rotation) of the end link in the root frame; q0 [B, dof] full qpos.
        Returns the target positions of the controlled joints [B, n_ctrl] (kinematics.py:243-260)."""
        kind = solver_config.get("type", "levenberg_marquardt")
        if getattr(self, "_ik_args", None) is not None and kind == "levenberg_marquardt":
            return self._compute_ik_kernel(delta_pose, float(solver_config.get("alpha", 1.0)))
        qc = q0[:, self.chain_dof_idx]
        _, _, J = self.chain.forward(qc)
        J = J[:, :, self.qmask]
        if solver_config.get("type", "levenberg_marquardt") == "levenberg_marquardt":
            lambd = 0.0001  # regularisation so that J^T J is non-singular (kinematics.py:245)
            JT = J.transpose(1, 2)
            A = torch.bmm(JT, J) + lambd * torch.eye(J.shape[2], device=self.device, dtype=J.dtype)
            rhs = torch.bmm(JT, delta_pose.unsqueeze(-1))
            dq = torch.linalg.solve(A, rhs)
        elif solver_config["type"] == "pseudo_inverse":
            dq = torch.linalg.pinv(J) @ delta_pose.unsqueeze(-1)
        else:
            raise NotImplementedError(solver_config["type"])
        return qc[:, self.qmask] + solver_config.get("alpha", 1.0) * dq.squeeze(-1)

    def _compute_ik_kernel(self, delta_pose: torch.Tensor, alpha: float):
        """The same step through the C-ABI: reads the articulation's row of the world's qpos buffer in place."""
        w, art = self._ik_world, self._ik_art
        h = self._ik_handles.get(alpha)
        if h is None:
            h = self._ik_handles[alpha] = w.create_ik(*self._ik_args, lambd=1e-4, alpha=alpha)
        n_ctrl = int(self.qmask.sum())
        out = torch.empty((delta_pose.shape[0], n_ctrl), dtype=torch.float32, device=self.device)
        n_art = max(w.n_art, 1)
        qpos = w.qpos[art.art_index:]            # row env * n_art + art_index, as a pointer offset + a stride of n_art rows
        w.ik_step(h, delta_pose.to(torch.float32).contiguous(), qpos, n_art * w.max_dof, out)
        return out
