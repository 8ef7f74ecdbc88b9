#!/usr/bin/env python
"""bench.py -- env steps/sec of BaseEnv.step() (the reference's `gpu_sim.py` protocol,
mani_skill/examples/benchmarking/gpu_sim.py:91-108: random actions in [-1, 1], fps = steps * num_envs / time).

    python bench.py --gpus 1 --steps 200 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --steps K --warmup W      # CPU restatement of the reference path on the host cores

A "step" is one control step = `sim_freq/control_freq` (5) fused physics substeps + evaluate + obs + reward of
PickCube-v1 at num_envs=4096 per GPU, obs_mode=state (BASELINE.json configs[1]).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes: physics 3696 B per env-substep x 5 substeps + 213 B action/obs I/O per env-step
BYTES_PHYSICS_PER_ENV_SUBSTEP = 3696
BYTES_IO_PER_ENV_STEP = 213


def usable_cores():
    """Host cores this process may actually use: min(os.cpu_count, scheduler affinity, cgroup cpu quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.samples = []
        self.stop = False
        self.index = index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        import statistics
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle)
def cpu_oracle_throughput(n_envs, control_steps, substeps=5, seed=0):
    """Times the CPU oracle (float32 build) on all host cores: one worker PROCESS per core, each stepping its share of
    the sample (the reference vectorises its CPU backend the same way, one env process per core,
    mani_skill/examples/benchmarking/gpu_sim.py:72-84).  Returns (env-steps/s, cores, seconds)."""
    from oracle import oracle as _o
    _o.build()
    cores = usable_cores()
    per = max(n_envs // cores, 1)
    start_at = time.time() + 8.0 + 0.05 * per + 0.05 * cores  # imports + world construction + warm-up happen before this instant
    env_ = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "oracle", "cpu_worker.py"), str(per), str(control_steps), str(seed + i),
                               repr(start_at)], stdout=subprocess.PIPE, text=True, env=env_) for i in range(cores)]
    times = [float(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    dt = max(times)
    return per * cores * control_steps / dt, cores, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot be installed here (sapien/PhysX absent, see
    DESIGN.md), so the CPU restatement (oracle/, kind="port") is timed on the host cores with every thread it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    sample_envs = 256 * cores
    # warm-up + K steps, each step = one control step of the bounded sample
    v, cores, dt = cpu_oracle_throughput(sample_envs, max(args.steps, 1))
    line = {
        "impl": "reference", "metric": "env steps/sec (PickCube-v1, state obs)", "value": v, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PickCube-v1 num_envs=4096/GPU state-only (configs[1]); CPU arm runs a bounded sample",
                   "sample_envs": sample_envs, "substeps_per_step": 5},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_envs} envs x {args.steps} control steps (5 substeps each), CPU oracle f32, {cores} worker processes"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import maniskill_b200 as ms
    from maniskill_b200.backend import BUF_ALL

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_envs = args.num_envs
    env = ms.make("PickCube-v1", num_envs=n_envs, obs_mode="state", device=dev)
    venv = ms.ManiSkillVectorEnv(env, auto_reset=not args.no_auto_reset)
    world = env.scene.world
    A = env.action_dim
    from maniskill_b200.dist import ObsGather, shard_seeds
    obs, _ = venv.reset(seed=shard_seeds(2022, n_envs * world_size, rank, world_size))  # seeds keep the global env id
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gather_buf = ObsGather(n_envs, obs.shape[1], obs.dtype, dev) if (args.gather_obs and world_size > 1) else None

    def one_step(actions):
        o, r, te, tr, info = venv.step(actions)
        if gather_buf is not None:
            gather_buf(o)
        return o, r, te, tr

    # ---------------- warm-up
    for _ in range(max(args.warmup, 3)):
        one_step(2 * torch.rand((n_envs, A), device=dev, generator=gen) - 1)
    barrier()
    # ---------------- timed: device-resident inputs ("value")
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    actions_all = 2 * torch.rand((args.steps, n_envs, A), device=dev, generator=gen) - 1
    launches0 = world.kernel_launches
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        torch.cuda.nvtx.range_push("timed")  # ncu --nvtx --nvtx-include "timed/" captures exactly these launches
        for i in range(args.steps):
            flush.fill_(float(i))  # evict L2 between timed iterations
            ev0[i].record()
            one_step(actions_all[i])
            ev1[i].record()
        torch.cuda.nvtx.range_pop()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = world.kernel_launches - launches0
    step_ms = sorted(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t_dev = sum(step_ms) / 1e3
    # ---------------- timed: end to end through the public API with host buffers ("e2e")
    h_actions = torch.empty((n_envs, A), dtype=torch.float32).pin_memory()
    h_obs = torch.empty(tuple(obs.shape), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((n_envs,), dtype=torch.float32).pin_memory()
    h_done = torch.empty((n_envs, 2), dtype=torch.bool).pin_memory()
    cpu_actions = (2 * torch.rand((args.steps, n_envs, A)) - 1)
    barrier()
    te0 = time.perf_counter()
    for i in range(args.steps):
        h_actions.copy_(cpu_actions[i])
        a = h_actions.to(dev, non_blocking=True)
        o, r, te, tr = one_step(a)
        h_obs.copy_(o, non_blocking=True)
        h_rew.copy_(r, non_blocking=True)
        h_done[:, 0].copy_(te, non_blocking=True)
        h_done[:, 1].copy_(tr, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    t_e2e = time.perf_counter() - te0
    # ---------------- physics kernel alone (roofline numerator): CUDA events around b2s_step on its stream
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    k_iters = 20
    torch.cuda.synchronize()
    kt = 0.0
    for i in range(k_iters):
        world.target_qpos[:, :7] = world.qpos[:, :7] + 0.1 * (2 * torch.rand((n_envs, 7), device=dev, generator=gen) - 1)
        world.apply(1 << 5)
        flush.fill_(1.0)
        kev[0].record()
        world.step(env._sim_steps_per_control, BUF_ALL)
        kev[1].record()
        torch.cuda.synchronize()
        kt += kev[0].elapsed_time(kev[1])
    k_ms = kt / k_iters
    # ---------------- max over ranks
    t = torch.tensor([t_dev, t_e2e, t_wall], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, t_wall = [float(x) for x in t.cpu()]
    total_envs = n_envs * world_size
    value = total_envs * args.steps / t_dev
    e2e = total_envs * args.steps / t_e2e
    peak, peak_src = read_peaks()
    substeps = env._sim_steps_per_control
    bytes_per_launch = n_envs * (BYTES_PHYSICS_PER_ENV_SUBSTEP * substeps)
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    line = {
        "metric": "env steps/sec (PickCube-v1, state obs)", "value": value, "unit": "env-steps/s", "n_gpus": world_size,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PickCube-v1 num_envs={n_envs}/GPU state-only, sim_freq=100 control_freq=20 (5 substeps/step), "
                               "15 position + 1 velocity iterations, pd_joint_delta_pos, auto-reset on (BASELINE.json configs[1])",
                   "num_envs_total": total_envs, "substeps_per_s": value * substeps, "l2": "256 MiB write between timed steps",
                   "obs_all_gather": bool(gather_buf is not None), "wall_s": t_wall,
                   "step_ms_median_rank0": step_ms[len(step_ms) // 2], "step_ms_max_rank0": step_ms[-1]},
        "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": n_envs * A * 4,
                "d2h_bytes_per_step": int(h_obs.numel() * 4 + h_rew.numel() * 4 + h_done.numel())},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": args.traffic_bytes, "peak_source": peak_src, "kernel": "b2s_step: 5 substeps x (kin + collide + manifest + rowfill + solve kernels) + fetch_kernel",
                     "kernel_ms": k_ms, "algorithmic_bytes_per_launch": bytes_per_launch},
    }
    # ---------------- supplementary: the RGBD half of BASELINE.json's metric (PickCube-v1 state+rgb+depth, one 128x128 camera)
    extra = None
    if not args.no_rgbd:
        env.close()
        n_v = args.num_envs
        env_v = ms.make("PickCube-v1", num_envs=n_v, obs_mode="state+rgb+depth", device=dev)
        venv_v = ms.ManiSkillVectorEnv(env_v)
        venv_v.reset(seed=shard_seeds(2022, n_v * world_size, rank, world_size))
        for _ in range(3):
            venv_v.step(2 * torch.rand((n_v, A), device=dev, generator=gen) - 1)
        barrier()
        kv = 20
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(kv):
            venv_v.step(2 * torch.rand((n_v, A), device=dev, generator=gen) - 1)
        r1.record()
        barrier()
        t_v = torch.tensor([r0.elapsed_time(r1) / 1e3], dtype=torch.float64, device=dev)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(kv):
            env_v._sensors.capture()
        q1.record()
        torch.cuda.synchronize()
        raster_ms = q0.elapsed_time(q1) / kv
        if world_size > 1:
            dist.all_reduce(t_v, op=dist.ReduceOp.MAX)
        px_bytes = n_v * 128 * 128 * 12  # Color rgba8 + PositionSegmentation 4 x int16 actually written per image
        extra = {"workload": f"PickCube-v1 num_envs={n_v}/GPU obs_mode=state+rgb+depth (1 camera 128x128), {kv} steps, no L2 flush",
                 "env_steps_per_s": n_v * world_size * kv / float(t_v.item()), "ms_per_step": float(t_v.item()) / kv * 1e3,
                 "raster_kernel_ms": raster_ms, "raster_write_GBps": px_bytes / (raster_ms * 1e-3) / 1e9,
                 "raster_frac_of_hbm_peak": px_bytes / (raster_ms * 1e-3) / 1e9 / peak}
        env = env_v
    line["state_rgbd"] = extra
    if rank == 0:
        if not args.no_cpu_baseline and world_size == 1:
            cores = usable_cores()
            v, cores, dt = cpu_oracle_throughput(256 * cores, 100)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{256 * cores} envs x 100 control steps (5 substeps each) of the same workload, CPU oracle f32, "
                                              f"{cores} worker processes (usable cores), {dt:.1f}s wall"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    env.close()
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--num-envs", type=int, default=4096, help="sub-scenes per GPU")
    ap.add_argument("--gather-obs", action="store_true", help="all-gather the flattened observation across ranks (NCCL)")
    ap.add_argument("--no-auto-reset", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rgbd", action="store_true", help="skip the supplementary state+RGBD measurement")
    ap.add_argument("--traffic-bytes", type=float, default=110.1e6,
                    help="dram bytes read + written by one b2s_step (5 substeps): sum over its six kernels of dram__bytes_read.sum + "
                         "dram__bytes_write.sum from the committed per-kernel `ncu --set full` captures (cold caches per kernel, i.e. an upper "
                         "bound for the graph replay where the exchange buffers stay in L2): profiles/r01_pipeline_kernels_ncu_summary.md")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
