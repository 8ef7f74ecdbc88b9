"""Times env.step with visual observations + the raster kernel alone: python tools/time_render.py [task] [obs_mode] [num_envs]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maniskill_b200 as ms
task = sys.argv[1] if len(sys.argv) > 1 else "PickCube-v1"
mode = sys.argv[2] if len(sys.argv) > 2 else "rgbd"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
env = ms.make(task, num_envs=N, obs_mode=mode)
venv = ms.ManiSkillVectorEnv(env)
venv.reset(seed=0)
dev = env.device
for _ in range(3): venv.step(2 * torch.rand((N, 8), device=dev) - 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for _ in range(iters): venv.step(2 * torch.rand((N, 8), device=dev) - 1)
e1.record(); torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / iters
e0.record()
for _ in range(iters): env._sensors.capture()
e1.record(); torch.cuda.synchronize()
ms_r = e0.elapsed_time(e1) / iters
ncam = len(env._sensors.cams)
px = sum(c["width"] * c["height"] for c in env._sensors.cams)
print(f"{task} {mode} N={N}: step {ms_step:.3f} ms -> {N/ms_step*1e3:.0f} env-steps/s ; raster alone {ms_r:.3f} ms ({ncam} cams) -> {N*px*12/ms_r/1e6:.1f} GB/s of render-target writes")
