"""TEST DOUBLE: maniskill_b200.backend.World's interface on top of the host emulation (tests/emu), with CPU torch
tensors aliasing the emulated buffers.  Lets the host logic (BaseEnv, controllers, obs/reward, partial reset) be
tested on a machine without a GPU; the product never uses it."""
import numpy as np
import torch

from emu import BUF_ALL, EmuWorld


class EmuBackendWorld:
    def __init__(self, cm):
        self.cm = cm
        self._w = EmuWorld(cm)
        w = self._w
        self.device = torch.device("cpu")
        self.n_envs, self.n_rows, self.n_link = w.n_envs, w.n_rows, w.n_link
        s = cm.scalars
        self.n_art, self.n_fb = s["n_art"], s["n_fb"]
        self.max_dof = max(s["max_dof_per_art"], 1)
        self.rigid_body_data = torch.from_numpy(w.rigid_body_data.reshape(-1, 13))
        self.qpos, self.qvel, self.qacc, self.qf = [torch.from_numpy(a) for a in (w.qpos, w.qvel, w.qacc, w.qf)]
        self.target_qpos, self.target_qvel = torch.from_numpy(w.target_qpos), torch.from_numpy(w.target_qvel)
        self._queries = {}
        self.kernel_launches = 0

    def body_view(self):
        return self.rigid_body_data.view(self.n_envs, self.n_rows, 13)

    def step(self, substeps=1, fetch_mask=0):
        self._w.step(substeps, fetch_mask)

    def apply(self, mask):
        self._w.apply(mask)

    def fetch(self, mask=BUF_ALL):
        self._w.fetch(mask)

    def update_kinematics(self):
        self._w.fetch(1 << 8)

    def create_contact_query(self, row_pairs):
        key = tuple(map(tuple, row_pairs))
        self._queries[key] = list(key)
        return key

    def query_contact_impulses(self, key):
        out = np.stack([self._w.pair_impulse(a, b) for a, b in self._queries[key]], axis=1)
        return torch.from_numpy(out.astype(np.float32))

    @property
    def overflow_flag(self):
        return torch.tensor([self._w.overflow()])

    def close(self):
        pass

    # ---- rendering: the CPU raster oracle stands in for b2s_camera_group_create / b2s_render (tests may use oracle/)
    def create_camera_group(self, cameras, visuals, outputs=3):
        from maniskill_b200.backend import CameraGroup
        pix = sum(int(c["width"]) * int(c["height"]) for c in cameras)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt)
        g = CameraGroup(self, None, cameras, z((self.n_envs, pix, 4), torch.uint8), z((self.n_envs, pix, 4), torch.int16),
                        z((self.n_envs, pix, 3), torch.uint8) if outputs & 4 else None, z((self.n_envs, pix), torch.int16) if outputs & 8 else None,
                        z((self.n_envs, pix), torch.int16) if outputs & 16 else None)
        g._visuals = visuals
        return g

    def render(self, group, env_mask=None):
        from oracle import raster
        out = raster.render(group._visuals, group.cameras, self.body_view().numpy())
        for i, (color, posseg) in enumerate(out):
            a, b = int(group._offsets[i]), int(group._offsets[i + 1])
            group._color[:, a:b] = torch.from_numpy(np.ascontiguousarray(color)).reshape(self.n_envs, -1, 4)
            group._posseg[:, a:b] = torch.from_numpy(np.ascontiguousarray(posseg)).reshape(self.n_envs, -1, 4)
        # the compact textures follow from the raw targets (render/shaders.py:74-83 texture_transforms)
        if group._rgb is not None:
            group._rgb[:] = group._color[..., :3]
        if group._depth is not None:
            group._depth[:] = -group._posseg[..., 2]
        if group._seg is not None:
            group._seg[:] = group._posseg[..., 3]
