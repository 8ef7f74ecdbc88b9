import sys; sys.path.insert(0, "/root/repo")
import torch, maniskill_b200 as ms
for task, cm in [("PushCube-v1", "pd_ee_target_delta_pos"), ("StackCube-v1", None), ("PullCube-v1", None),
                 ("LiftPegUpright-v1", None), ("PokeCube-v1", None), ("RollBall-v1", None), ("PlaceSphere-v1", None), ("StackPyramid-v1", None), ("PullCubeTool-v1", None), ("PlugCharger-v1", None)]:
    env = ms.ManiSkillVectorEnv(ms.make(task, num_envs=1024, obs_mode="state", control_mode=cm), auto_reset=True)
    obs, _ = env.reset(seed=0)
    A = env.base_env.action_dim
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(25):
        if i == 5: ev[0].record()
        obs, r, te, tr, info = env.step(2 * torch.rand((1024, A), device="cuda", generator=g) - 1)
    ev[1].record(); torch.cuda.synchronize()
    print(task, tuple(obs.shape), "finite", bool(torch.isfinite(obs).all() and torch.isfinite(r).all()), f"{ev[0].elapsed_time(ev[1])/20:.2f} ms/step", "overflow", int(env.base_env.scene.world.overflow_flag.item()), flush=True)
