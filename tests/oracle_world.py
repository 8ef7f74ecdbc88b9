"""TEST DOUBLE: maniskill_b200.backend.World's interface on top of the CPU oracle (oracle/), CPU torch tensors as the exposed buffers.

Used where the reference's `sim_backend="physx_cpu"` path is exercised through the shim (tests/cpu_sim_double.py): BASELINE.json configs[0]
(PickCube-v1, num_envs=1, CPU simulation) and the reference's CPU-vs-GPU tests, with the oracle standing where SAPIEN's CPU PhysX stands.  The product never
imports this file (nor oracle/): `sapien.physx.PhysxCpuSystem()` of the shim raises."""
import numpy as np
import torch

from maniskill_b200.backend import BUF_ALL, BUF_QF, BUF_QPOS, BUF_QVEL, BUF_RIGID, BUF_ROOT_POSE, BUF_TARGET_QPOS, BUF_TARGET_QVEL
from oracle.oracle import OracleWorld

_JOINT = (("qpos", BUF_QPOS), ("qvel", BUF_QVEL), ("qf", BUF_QF), ("target_qpos", BUF_TARGET_QPOS), ("target_qvel", BUF_TARGET_QVEL))


class OracleBackendWorld:
    def __init__(self, cm, precision="f32"):
        self.cm = cm
        self._o = OracleWorld(cm, precision)
        s = cm.scalars
        self.device = torch.device("cpu")
        self.n_envs, self.n_rows, self.n_link, self.n_art, self.n_fb = s["n_envs"], cm.n_rows, s["n_link"], s["n_art"], s["n_fb"]
        assert self.n_rows == self.n_link + self.n_fb
        self.max_dof = max(s["max_dof_per_art"], 1)
        N, na = self.n_envs, max(self.n_art, 1)
        self.rigid_body_data = torch.zeros((N * self.n_rows, 13), dtype=torch.float32)
        self.qpos, self.qvel, self.qacc, self.qf, self.target_qpos, self.target_qvel = [torch.zeros((N * na, self.max_dof), dtype=torch.float32) for _ in range(6)]
        names = sorted(cm.art_index, key=cm.art_index.get)
        self._dof = [(cm.art_dof_start[i], cm.art_dof_start[i] + len(cm.dof_names[n])) for i, n in enumerate(names)]
        self._root_rows = [cm.link_info[n][0].row for n in names]
        self.kernel_launches = 0
        self.fetch()

    def body_view(self):
        return self.rigid_body_data.view(self.n_envs, self.n_rows, 13)

    # exposed buffers -> oracle state
    def apply(self, mask=BUF_ALL):
        body = self.body_view().double().numpy()
        if mask & BUF_RIGID and self.n_fb:
            self._o.set_bodies(body[:, self.n_link:, :])
        if mask & BUF_ROOT_POSE and self.n_art:
            self._o.set_roots(body[:, self._root_rows, :7])
        for name, bit in _JOINT:
            if mask & bit and self.n_art:
                buf = getattr(self, name).view(self.n_envs, -1, self.max_dof).double().numpy()
                full = self._o.get_joint(name)
                for a, (lo, hi) in enumerate(self._dof):
                    full[:, lo:hi] = buf[:, a, :hi - lo]
                self._o.set_joint(name, full)

    # oracle state -> exposed buffers
    def fetch(self, mask=BUF_ALL):
        self.body_view()[:] = torch.from_numpy(self._o.rigid_body_data().astype(np.float32))
        if self.n_art:
            for name in ("qpos", "qvel", "qacc", "qf", "target_qpos", "target_qvel"):
                full = self._o.get_joint(name)
                buf = getattr(self, name).view(self.n_envs, -1, self.max_dof)
                for a, (lo, hi) in enumerate(self._dof):
                    buf[:, a, :hi - lo] = torch.from_numpy(full[:, lo:hi].astype(np.float32))

    def step(self, substeps=1, fetch_mask=BUF_ALL):
        self._o.step(substeps)
        self.fetch()

    def update_kinematics(self):
        self.fetch()

    def create_contact_query(self, row_pairs):
        return tuple(map(tuple, row_pairs))

    def query_contact_impulses(self, key):
        return torch.from_numpy(np.stack([self._o.pair_impulse(a, b) for a, b in key], axis=1).astype(np.float32))

    def contacts(self, env=0):
        """rows of (rowA, rowB, position 3, normal 3, separation, impulse 3) of the last substep"""
        return self._o.contacts(env)

    @property
    def overflow_flag(self):
        return torch.tensor([self._o.overflow()])

    def close(self):
        pass
