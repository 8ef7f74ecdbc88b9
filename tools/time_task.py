"""Device time per control step of any registered task (state observations, random actions): python tools/time_task.py <task> <num_envs> [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maniskill_b200 as ms
task, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
env = ms.ManiSkillVectorEnv(ms.make(task, num_envs=n, obs_mode="state"), auto_reset=True)
env.reset(seed=0)
A = env.base_env.action_dim
g = torch.Generator(device="cuda"); g.manual_seed(0)
for _ in range(10):
    env.step(2 * torch.rand((n, A), device="cuda", generator=g) - 1)
acts = 2 * torch.rand((steps, n, A), device="cuda", generator=g) - 1
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    env.step(acts[i])
e1.record(); torch.cuda.synchronize()
w = env.base_env.scene.world
ms_step = e0.elapsed_time(e1) / steps
e0.record()
from maniskill_b200.backend import BUF_ALL
for _ in range(20): w.step(5, BUF_ALL)
e1.record(); torch.cuda.synchronize()
print(f"{task} N={n} KIN_GROUP={os.environ.get('B2S_KIN_GROUP','0')}: {ms_step:.3f} ms/step -> {n/ms_step*1e3:.0f} env-steps/s ; b2s_step alone {e0.elapsed_time(e1)/20:.3f} ms ; overflow {int(w.overflow_flag.item())}")
