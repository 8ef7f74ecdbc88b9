import sys, os
sys.path.insert(0, "/root/repo")
import torch
import maniskill_b200 as ms
from maniskill_b200.backend import BUF_ALL
N=4096
env = ms.make("PickCube-v1", num_envs=N, obs_mode="state")
env.reset(seed=0)
for _ in range(5): env.step(2*torch.rand((N,8),device=env.device)-1)
w = env.scene.world
torch.cuda.synchronize()
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e0.record()
    for _ in range(20): w.step(5, BUF_ALL)
    e1.record(); torch.cuda.synchronize()
    print("B2S_PDL=%s b2s_step %.4f ms" % (os.environ.get("B2S_PDL","0"), e0.elapsed_time(e1)/20))
