"""CPU: the N>1 path (env sharding + observation all-gather) with world_size-2 gloo processes on the emulated backend."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maniskill_b200.dist import shard_range, shard_seeds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for total in (7, 8, 4096, 16384):
        for g in (1, 2, 3, 8):
            blocks = [shard_range(total, r, g) for r in range(g)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(g - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard_seeds(2022, 8, 1, 2) == [2026, 2027, 2028, 2029]
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world_size, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import maniskill_b200 as ms
    from emu_world import EmuBackendWorld
    from maniskill_b200.dist import ObsGather
    n_local = 2
    env = ms.make("PickCube-v1", num_envs=n_local, obs_mode="state", world_factory=EmuBackendWorld)
    obs, _ = env.reset(seed=shard_seeds(2022, n_local * world_size, rank, world_size))
    gather = ObsGather(n_local, obs.shape[1])
    a = torch.full((n_local, 8), 0.1 * (rank + 1))
    obs, rew, *_ = env.step(a)
    full = gather(obs).clone()
    # every rank must hold every rank's block, in rank order
    mine = full[rank * n_local:(rank + 1) * n_local]
    ok = torch.equal(mine, obs) and full.shape == (n_local * world_size, obs.shape[1])
    blocks = [torch.empty_like(obs) for _ in range(world_size)]
    dist.all_gather(blocks, obs)
    ok = ok and torch.equal(torch.cat(blocks), full)
    q.put((rank, bool(ok), float(full.abs().sum())))
    dist.destroy_process_group()


def test_two_rank_gloo_obs_all_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert abs(res[0][2] - res[1][2]) < 1e-3  # both ranks see the same gathered batch


REF = "/root/reference"


def _reference_worker(rank, world_size, port, q):
    """one process per GPU, each running the UNMODIFIED reference with num_envs = N / G on its own world (here: the emulated one), seeds by global
    sub-scene id, the flattened state observation all-gathered."""
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import maniskill_b200.compat as compat
    from emu_world import EmuBackendWorld
    from maniskill_b200.dist import ObsGather
    compat.install()
    compat.WORLD_FACTORY = lambda cm, dev: EmuBackendWorld(cm)
    sys.path.insert(0, REF)
    import gymnasium as gym
    import mani_skill.envs  # noqa: F401
    import mani_skill.envs.sapien_env as SE
    import mani_skill.envs.utils.system.backend as B
    orig = B.parse_sim_and_render_backend

    def parse(sim_backend, render_backend):
        info = orig(sim_backend, render_backend)
        info.device = torch.device("cpu")
        return info
    SE.parse_sim_and_render_backend = parse
    n_local = 2
    env = gym.make("PickCube-v1", num_envs=n_local, obs_mode="state", sim_backend=f"physx_cuda:{rank}")
    obs, _ = env.reset(seed=shard_seeds(2022, n_local * world_size, rank, world_size))
    gather = ObsGather(n_local, obs.shape[1])
    obs, rew, *_ = env.step(torch.full((n_local, 8), 0.1 * (rank + 1)))
    full = gather(obs).clone()
    ok = torch.equal(full[rank * n_local:(rank + 1) * n_local], obs) and full.shape == (n_local * world_size, 42)
    # the robot's initial joint noise comes from the per-sub-scene numpy generators (seed = global id): the two ranks drew different sub-scenes
    other = full[(1 - rank) * n_local:(2 - rank) * n_local]
    ok = ok and float((other[:, :7] - obs[:, :7]).abs().max()) > 1e-4
    q.put((rank, bool(ok), float(full.abs().sum())))
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mani_skill")), reason="needs the reference checkout at /root/reference")
def test_two_ranks_of_the_unmodified_reference():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_reference_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert abs(res[0][2] - res[1][2]) < 1e-3
