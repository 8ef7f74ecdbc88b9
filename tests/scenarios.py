"""Shared seeded scenarios for the parity tests (oracle vs host emulation vs CUDA through the C-ABI)."""
import numpy as np

from maniskill_b200.scenes import PANDA_REST_QPOS, pick_cube_scene


def pick_cube_random_actions(n_envs, n_substeps=100, seed=0, action_every=5):
    """Yields (step, target_qpos or None). PickCube reset state + random joint-delta targets, like the benchmark's
    random actions (mani_skill/examples/benchmarking/gpu_sim.py:97-108)."""
    rng = np.random.RandomState(seed)
    q0 = PANDA_REST_QPOS + rng.normal(0, 0.02, (n_envs, 9))
    q0[:, 7:] = 0.04
    cube = np.zeros((n_envs, 13))
    cube[:, 0:2] = rng.uniform(-0.1, 0.1, (n_envs, 2))
    cube[:, 2] = 0.02
    yaw = rng.uniform(0, 2 * np.pi, n_envs)
    cube[:, 3] = np.cos(yaw / 2)
    cube[:, 6] = np.sin(yaw / 2)
    deltas = [rng.uniform(-0.1, 0.1, (n_envs, 9)) for _ in range(n_substeps // action_every + 1)]
    grip = [rng.uniform(-0.01, 0.04, (n_envs, 1)) for _ in range(n_substeps // action_every + 1)]
    return q0, cube, deltas, grip


def run_pick_cube(world_kind, cm, n_substeps=100, seed=0, device=None, emu_mode=1):
    """Runs the scenario on 'oracle32' / 'oracle64' / 'emu' / 'cuda'; returns dict of final numpy arrays."""
    N = cm.scalars["n_envs"]
    q0, cube, deltas, grip = pick_cube_random_actions(N, n_substeps, seed)
    cube_fb = cm.actor_fb["cube"]
    n_link = cm.scalars["n_link"]
    if world_kind.startswith("oracle"):
        from oracle.oracle import OracleWorld
        w = OracleWorld(cm, "f32" if world_kind == "oracle32" else "f64")
        w.set_joint("qpos", q0)
        w.set_joint("target_qpos", q0)
        b = w.get_bodies()
        b[:, cube_fb] = cube
        w.set_bodies(b)
        for s in range(n_substeps):
            if s % 5 == 0:
                tq = w.get_joint("qpos") + deltas[s // 5]
                tq[:, 7:] = grip[s // 5]
                w.set_joint("target_qpos", tq)
            w.step(1)
        return dict(qpos=w.get_joint("qpos"), qvel=w.get_joint("qvel"), body=w.rigid_body_data(), world=w)
    if world_kind == "emu":
        from emu import EmuWorld
        w = EmuWorld(cm)
        w.split = bool(emu_mode)  # 0 single-lane fused substep, 1 pipelined substep (the CUDA library's default)
        w.qpos[:] = q0
        w.target_qpos[:] = q0
        w.rigid_body_data[:, n_link + cube_fb] = cube
        w.apply()
        for s in range(n_substeps):
            if s % 5 == 0:
                w.fetch(4)
                tq = w.qpos.astype(np.float64) + deltas[s // 5]
                tq[:, 7:] = grip[s // 5]
                w.target_qpos[:] = tq
                w.apply(32)
            w.step(1, 0)
        w.fetch()
        return dict(qpos=w.qpos.astype(np.float64), qvel=w.qvel.astype(np.float64), body=w.rigid_body_data.astype(np.float64), world=w)
    if world_kind == "cuda":
        import torch
        from maniskill_b200.backend import BUF_ALL, World
        w = World(cm, device)
        dev = w.device
        w.qpos[:] = torch.tensor(q0, dtype=torch.float32, device=dev)
        w.target_qpos[:] = w.qpos
        w.body_view()[:, n_link + cube_fb] = torch.tensor(cube, dtype=torch.float32, device=dev)
        w.apply()
        for s in range(n_substeps):
            if s % 5 == 0:
                w.fetch(4)
                tq = w.qpos.double() + torch.tensor(deltas[s // 5], device=dev)
                tq[:, 7:] = torch.tensor(grip[s // 5], device=dev)
                w.target_qpos[:] = tq.float()
                w.apply(32)
            w.step(1, 0)
        w.fetch(BUF_ALL)
        torch.cuda.synchronize()
        return dict(qpos=w.qpos.double().cpu().numpy(), qvel=w.qvel.double().cpu().numpy(),
                    body=w.body_view().double().cpu().numpy(), world=w)
    raise ValueError(world_kind)


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor), elementwise"""
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def rel_err_vec(a, b, floor):
    """max over vectors (last axis) of ||a-b|| / max(||b||, floor): error relative to the magnitude of the quantity
    (a position, a unit quaternion, a joint vector), not to each component -- a 1e-7 absolute wobble of a quaternion
    component that happens to be 1e-3 is float32 rounding, not a 1e-4 relative error of the pose."""
    d = np.linalg.norm(a - b, axis=-1)
    n = np.maximum(np.linalg.norm(b, axis=-1), floor)
    return float(np.max(d / n))


def run_task_vs_oracle(task, steps, n, world_factory=None, seed=4):
    """Steps a registered task with seeded random actions on the product backend (CUDA through the C-ABI; or the host emulation when
    `world_factory` is given) and the CPU oracle started from the same state and fed the same drive targets.  Returns
    (max |qpos error|, max |position error|, overflow flag of the backend)."""
    import torch

    import maniskill_b200 as ms
    from oracle.oracle import OracleWorld
    kw = dict(device="cpu", world_factory=world_factory) if world_factory is not None else {}
    env = ms.make(task, num_envs=n, obs_mode="state", **kw)
    obs, _ = env.reset(seed=seed)
    w, cm = env.scene.world, env.cm
    n_art, n_link = cm.scalars["n_art"], cm.scalars["n_link"]
    o = OracleWorld(cm, "f32")

    def joint_state(t):  # exposed [N*n_art, max_dof] -> oracle [N, n_dof]
        t = t.double().cpu().numpy().reshape(n, n_art, -1)
        return np.concatenate([t[:, a, :cm.art_dof_start[a + 1] - cm.art_dof_start[a]] for a in range(n_art)], axis=1)

    for name, buf in (("qpos", w.qpos), ("qvel", w.qvel), ("target_qpos", w.target_qpos), ("target_qvel", w.target_qvel)):
        o.set_joint(name, joint_state(buf))
    body = w.body_view().double().cpu().numpy()
    o.set_bodies(body[:, n_link:])
    o.set_roots(np.stack([body[:, cm.arrays["art_link_start"][a], :7] for a in range(n_art)], axis=1))
    g = torch.Generator(device=env.device).manual_seed(0)
    for _ in range(steps):
        a = 2 * torch.rand((n, env.action_dim), device=env.device, generator=g) - 1
        obs, rew, term, trunc, info = env.step(a)
        o.set_joint("target_qpos", joint_state(w.target_qpos))
        o.set_joint("target_qvel", joint_state(w.target_qvel))
        if task == "OpenCabinetDrawer-v1":  # the goal marker is re-posed by the task every step
            o.set_bodies(w.body_view().double().cpu().numpy()[:, n_link:])
        o.step(5)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    err_q = float(np.abs(joint_state(w.qpos) - o.get_joint("qpos")).max())
    err_p = float(np.abs(w.body_view().double().cpu().numpy()[..., :3] - o.rigid_body_data()[..., :3]).max())
    overflow = int(w.overflow_flag.item())
    env.close()
    return err_q, err_p, overflow
