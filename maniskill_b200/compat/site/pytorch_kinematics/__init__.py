"""The part of `pytorch_kinematics` the reference's end-effector controllers use (mani_skill/agents/controllers/utils/kinematics.py:
142-260): a serial chain built from a URDF string, batched forward kinematics and the geometric Jacobian in the base frame.

    chain = build_serial_chain_from_urdf(urdf_bytes, end_link_name).to(device=...)
    chain.forward_kinematics(q).get_matrix()     # [B, 4, 4] end-link pose in the root frame
    chain.jacobian(q)                            # [B, 6, n]  rows: linear velocity (3), angular velocity (3) of the end link
    chain.get_joint_limits()                     # (lower[n], upper[n])
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import List, Optional

import numpy as np
import torch


def _rpy_to_mat(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class Transform3d:
    def __init__(self, matrix: torch.Tensor):
        self._m = matrix

    def get_matrix(self) -> torch.Tensor:
        return self._m


class SerialChain:
    def __init__(self, joints: List[dict], dtype=torch.float32, device="cpu"):
        self._joints = joints      # dict(name, type, origin 4x4, axis 3, lower, upper), root -> end
        self.dtype, self.device = dtype, torch.device(device)
        self._build()

    def _build(self):
        mov = [j for j in self._joints if j["type"] != "fixed"]
        self.n_joints = len(mov)
        self._origin = [torch.as_tensor(j["origin"], dtype=self.dtype, device=self.device) for j in self._joints]
        self._axis = [torch.as_tensor(j["axis"], dtype=self.dtype, device=self.device) for j in self._joints]

    def to(self, dtype=None, device=None):
        self.dtype = dtype or self.dtype
        self.device = torch.device(device) if device is not None else self.device
        self._build()
        return self

    def get_joint_parameter_names(self, exclude_fixed=True):
        return [j["name"] for j in self._joints if not (exclude_fixed and j["type"] == "fixed")]

    def get_joint_limits(self):
        mov = [j for j in self._joints if j["type"] != "fixed"]
        return [j["lower"] for j in mov], [j["upper"] for j in mov]

    def _walk(self, th: torch.Tensor):
        """-> end transform [B,4,4], per moving joint: world axis [B,3], world origin [B,3], is_revolute."""
        th = torch.as_tensor(th, dtype=self.dtype, device=self.device)
        if th.ndim == 1:
            th = th[None]
        B = th.shape[0]
        T = torch.eye(4, dtype=self.dtype, device=self.device).expand(B, 4, 4).clone()
        axes, origins, kinds = [], [], []
        k = 0
        for j, O, ax in zip(self._joints, self._origin, self._axis):
            T = T @ O
            if j["type"] == "fixed":
                continue
            q = th[:, k]
            k += 1
            axes.append(T[:, :3, :3] @ ax)
            origins.append(T[:, :3, 3])
            M = torch.eye(4, dtype=self.dtype, device=self.device).expand(B, 4, 4).clone()
            if j["type"] == "prismatic":
                kinds.append(False)
                M[:, :3, 3] = ax[None] * q[:, None]
            else:
                kinds.append(True)
                x, y, z = ax
                c, s = torch.cos(q), torch.sin(q)
                C = 1 - c
                M[:, 0, 0] = c + x * x * C; M[:, 0, 1] = x * y * C - z * s; M[:, 0, 2] = x * z * C + y * s
                M[:, 1, 0] = y * x * C + z * s; M[:, 1, 1] = c + y * y * C; M[:, 1, 2] = y * z * C - x * s
                M[:, 2, 0] = z * x * C - y * s; M[:, 2, 1] = z * y * C + x * s; M[:, 2, 2] = c + z * z * C
            T = T @ M
        return T, axes, origins, kinds

    def forward_kinematics(self, th, end_only: bool = True):
        return Transform3d(self._walk(th)[0])

    def jacobian(self, th, locations=None, ret_eef_pose: bool = False):
        T, axes, origins, kinds = self._walk(th)
        pe = T[:, :3, 3]
        cols = []
        for a, o, rev in zip(axes, origins, kinds):
            if rev:
                cols.append(torch.cat([torch.cross(a, pe - o, dim=1), a], dim=1))
            else:
                cols.append(torch.cat([a, torch.zeros_like(a)], dim=1))
        J = torch.stack(cols, dim=2) if cols else torch.zeros((T.shape[0], 6, 0), dtype=self.dtype, device=self.device)
        return (J, T) if ret_eef_pose else J


def build_serial_chain_from_urdf(data, end_link_name: str, root_link_name: str = "") -> SerialChain:
    if isinstance(data, bytes):
        data = data.decode("utf-8")
    root = ET.fromstring(data)
    parent_of = {}
    for j in root.findall("joint"):
        parent_of[j.find("child").get("link")] = j
    chain, cur = [], end_link_name
    while cur in parent_of and cur != root_link_name:
        j = parent_of[cur]
        chain.append(j)
        cur = j.find("parent").get("link")
    joints = []
    for j in reversed(chain):
        o = j.find("origin")
        xyz = [float(v) for v in o.get("xyz", "0 0 0").split()] if o is not None else [0, 0, 0]
        rpy = [float(v) for v in o.get("rpy", "0 0 0").split()] if o is not None else [0, 0, 0]
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = _rpy_to_mat(*rpy), xyz
        ax = [float(v) for v in j.find("axis").get("xyz").split()] if j.find("axis") is not None else [1.0, 0, 0]
        n = np.linalg.norm(ax)
        ax = (np.asarray(ax) / n).tolist() if n > 0 else [1.0, 0, 0]
        lim = j.find("limit")
        jt = j.get("type")
        lo = float(lim.get("lower", -np.pi)) if lim is not None and jt != "continuous" else -np.pi
        hi = float(lim.get("upper", np.pi)) if lim is not None and jt != "continuous" else np.pi
        joints.append(dict(name=j.get("name"), type="fixed" if jt == "fixed" else ("prismatic" if jt == "prismatic" else "revolute"), origin=T, axis=ax,
                           lower=lo, upper=hi))
    return SerialChain(joints)


class PseudoInverseIK:
    """Constructed by the reference's Kinematics._setup_gpu (kinematics.py:160-166) but not used by its controllers' IK step."""

    def __init__(self, serial_chain, **kw):
        self.chain, self.config = serial_chain, kw

    def solve(self, *a, **kw):
        raise NotImplementedError("the iterative PseudoInverseIK solver is not part of this subset")
