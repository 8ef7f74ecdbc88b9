"""Times the fused substep kernel alone (CUDA events) for PickCube at several env counts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from maniskill_b200.backend import World, BUF_ALL
from maniskill_b200.scenes import pick_cube_scene, PANDA_REST_QPOS

for N in [int(a) for a in (sys.argv[1:] or ["1024", "4096", "16384"])]:
    cm = pick_cube_scene(N).compile()
    t0 = time.time()
    w = World(cm)
    dev = w.device
    rng = np.random.RandomState(0)
    q0 = PANDA_REST_QPOS + rng.normal(0, 0.02, (N, 9)); q0[:, 7:] = 0.04
    w.qpos[:] = torch.tensor(q0, dtype=torch.float32, device=dev)
    w.target_qpos[:] = w.qpos
    w.apply()
    torch.cuda.synchronize()
    t_create = time.time() - t0
    for _ in range(3):
        w.step(5, BUF_ALL)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    iters = 60
    torch.manual_seed(0)
    ev[0].record()
    for i in range(iters):
        w.target_qpos[:] = w.qpos + 0.1 * (2 * torch.rand_like(w.qpos) - 1)
        w.apply(32)
        w.step(5, BUF_ALL)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / iters
    print(f"N={N}: create {t_create:.2f}s, control step (5 substeps) {ms:.3f} ms -> {N/ms*1e3:.0f} env-steps/s, overflow={int(w.overflow_flag.item())}", flush=True)
    w.close()
