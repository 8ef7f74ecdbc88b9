"""Short rollout of any registered task for an ncu launch list: python tools/prof_task.py <task> <num_envs> <obs_mode> <steps>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maniskill_b200 as ms

task, n, mode, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
env = ms.ManiSkillVectorEnv(ms.make(task, num_envs=n, obs_mode=mode), auto_reset=True)
env.reset(seed=0)
A = env.base_env.action_dim
g = torch.Generator(device="cuda")
g.manual_seed(0)
for _ in range(20):
    env.step(2 * torch.rand((n, A), device="cuda", generator=g) - 1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
for _ in range(steps):
    env.step(2 * torch.rand((n, A), device="cuda", generator=g) - 1)
torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
