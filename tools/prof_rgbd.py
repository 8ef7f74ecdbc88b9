"""Short state+rgbd rollout for an ncu launch list: python tools/prof_rgbd.py [num_envs] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maniskill_b200 as ms

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
env = ms.make("PickCube-v1", num_envs=N, obs_mode="state+rgb+depth")
env = ms.ManiSkillVectorEnv(env, auto_reset=True)
env.reset(seed=0)
g = torch.Generator(device="cuda")
g.manual_seed(0)
for i in range(3):
    env.step(2 * torch.rand((N, env.action_dim if hasattr(env, "action_dim") else 8), device="cuda", generator=g) - 1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(steps):
    env.step(2 * torch.rand((N, 8), device="cuda", generator=g) - 1)
ev[1].record()
torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print(f"{ev[0].elapsed_time(ev[1]) / steps:.3f} ms per step")
