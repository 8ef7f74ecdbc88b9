"""Env-steps/s of the other registered tasks (device-timed, random actions, auto-reset): python tools/bench_tasks.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maniskill_b200 as ms

CASES = [("PickCube-v1", 4096, "state"), ("PickCube-v1", 16384, "state"), ("PickCube-v1", 4096, "state+rgb+depth"),
         ("PegInsertionSide-v1", 4096, "state"), ("PegInsertionSide-v1", 4096, "rgbd"), ("OpenCabinetDrawer-v1", 2048, "state")]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]
for task, n, mode in CASES:
    env = ms.ManiSkillVectorEnv(ms.make(task, num_envs=n, obs_mode=mode), auto_reset=True)
    env.reset(seed=0)
    A = env.base_env.action_dim
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    steps = 60 if "rgb" not in mode else 20
    for _ in range(5):
        env.step(2 * torch.rand((n, A), device="cuda", generator=g) - 1)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        env.step(2 * torch.rand((n, A), device="cuda", generator=g) - 1)
    ev[1].record()
    torch.cuda.synchronize()
    ms_step = ev[0].elapsed_time(ev[1]) / steps
    print(f"{task} {mode} num_envs={n}: {ms_step:.3f} ms/step -> {n / ms_step * 1e3:,.0f} env-steps/s "
          f"(overflow flag {int(env.base_env.scene.world.overflow_flag.item())})", flush=True)
    env.close()
