"""One control step's kernels inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`): the b2s_step graph replay
(5 substeps x 6 kernels + fetch) and one raster launch, after warm-up, caches as the running benchmark leaves them.

    ncu --profile-from-start off --cache-control none --clock-control none --csv --log-file gpurun_out/r02_traffic.csv \\
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum python tools/prof_traffic.py
    python tools/ncu_traffic.py gpurun_out/r02_traffic.csv          # -> profiles/r02_traffic.json (+ a per-kernel table)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maniskill_b200 as ms
from maniskill_b200.backend import BUF_ALL

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = ms.make("PickCube-v1", num_envs=N, obs_mode="state+rgb+depth")
venv = ms.ManiSkillVectorEnv(env)
venv.reset(seed=0)
for _ in range(8):
    venv.step(2 * torch.rand((N, 8), device=env.device) - 1)
torch.cuda.synchronize()
w = env.scene.world
torch.cuda.profiler.start()
w.step(5, BUF_ALL)
env._sensors.capture()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
