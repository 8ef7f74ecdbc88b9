"""Short run for ncu: PickCube N envs, a few control steps of random actions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maniskill_b200.backend import World, BUF_ALL
from maniskill_b200.scenes import pick_cube_scene, PANDA_REST_QPOS
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cm = pick_cube_scene(N).compile()
w = World(cm)
rng = np.random.RandomState(0)
q0 = PANDA_REST_QPOS + rng.normal(0, 0.02, (N, 9)); q0[:, 7:] = 0.04
w.qpos[:] = torch.tensor(q0, dtype=torch.float32, device=w.device)
w.target_qpos[:] = w.qpos
w.apply()
for i in range(steps):
    w.target_qpos[:] = w.qpos + 0.1 * (2 * torch.rand_like(w.qpos) - 1)
    w.apply(32)
    w.step(5, BUF_ALL)
torch.cuda.synchronize()
print("done", int(w.overflow_flag.item()))
