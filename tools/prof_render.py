"""Short run for ncu: a few raster launches (PickCube-v1 rgbd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maniskill_b200 as ms
task = sys.argv[1] if len(sys.argv) > 1 else "PickCube-v1"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = ms.make(task, num_envs=N, obs_mode="rgbd")
env.reset(seed=0)
for _ in range(3):
    env.step(2 * torch.rand((N, 8), device=env.device) - 1)
for _ in range(4):
    env._sensors.capture()
torch.cuda.synchronize()
print("done")
