"""b2s_step alone in the benchmark's steady state (random actions, auto-reset): python tools/time_b2s_step.py <task> <num_envs>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maniskill_b200 as ms
from maniskill_b200.backend import BUF_ALL
task, n = sys.argv[1], int(sys.argv[2])
env = ms.ManiSkillVectorEnv(ms.make(task, num_envs=n, obs_mode="state"), auto_reset=True)
env.reset(seed=0)
A = env.base_env.action_dim
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = env.base_env.scene.world
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot, cnt = 0.0, 0
for i in range(60):
    env.step(2 * torch.rand((n, A), device="cuda", generator=g) - 1)
    if i >= 20 and i % 4 == 0:   # a few replays from the live state (targets stay: the drives keep pulling)
        e0.record()
        for _ in range(4): w.step(5, BUF_ALL)
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) / 4; cnt += 1
print(f"{task} N={n} KIN_GROUP={os.environ.get('B2S_KIN_GROUP','0')}: b2s_step {tot/cnt:.3f} ms (live state) overflow {int(w.overflow_flag.item())}")
