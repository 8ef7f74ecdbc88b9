#!/usr/bin/env python
"""bench.py -- env steps/sec of BaseEnv.step() (the reference's `gpu_sim.py` protocol,
mani_skill/examples/benchmarking/gpu_sim.py:91-108: random actions in [-1, 1], fps = steps * num_envs / time).

    python bench.py --gpus 1 --steps 100 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --steps K --warmup W      # CPU restatement of the reference path on the host cores

Headline = BASELINE.json's metric: PickCube-v1 at num_envs=4096 per GPU with obs_mode=state+rgb+depth (one 128x128 camera), stock
simulation config (5 substeps per control step, 15 position + 1 velocity iterations), pd_joint_delta_pos, auto-reset on.  A "step" is one
control step of every sub-scene: controller + 5 physics substeps + evaluate + reward + observation + camera render (+ the auto-reset of
finished sub-scenes).  The state-only figure (configs[1]) and the 1024 / 16384 sizes are side keys of the same JSON line (rank 0).
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

# algorithmic bytes: physics 3696 B per env-substep, 213 B action/obs I/O per env-step, delivered pixels
# rgb 3 B + depth 2 B per pixel of the 128 x 128 camera
BYTES_PHYSICS_PER_ENV_SUBSTEP = 3696
BYTES_IO_PER_ENV_STEP = 213
BYTES_PER_DELIVERED_PIXEL = 5
CAMERA_PIXELS = 128 * 128

WORKLOAD = ("PickCube-v1 num_envs={n}/GPU obs_mode={mode} (state 42 floats + one 128x128 camera: rgb uint8 x3 + depth int16), sim_freq=100 "
            "control_freq=20 (5 substeps/step), 15 position + 1 velocity iterations, pd_joint_delta_pos, auto-reset on (BASELINE.json metric)")


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


def read_traffic():
    """dram bytes per launch of the profiled kernels, parsed from the committed ncu summary (profiles/r02_traffic.json, written by
    tools/ncu_traffic.py from an `ncu --set full` capture); {} when there is none."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler:
    """SM clocks / throttle reasons of the job's GPUs sampled DURING the timed region, in-process through NVML (rank 0 only: one thread,
    no subprocesses -- eight ranks each forking nvidia-smi five times a second is a host-side disturbance of its own)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, indices):
        self.indices = list(indices)
        self.samples = []
        self.stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(phys[i] if phys else i) for i in self.indices]
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        nv = self.nvml
        for h in self.handles:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.samples.append((float(sm), float(mx), int(bits)))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", ",".join(map(str, self.indices))],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        for line in out.splitlines():
            s = [x.strip() for x in line.split(",")]
            bits = sum(b for (name, b), v in zip(self.REASONS, s[2:6]) if v.lower().startswith("active"))
            self.samples.append((float(s[0]), float(s[1]), bits))

    def _run(self):
        while not self.stop:
            try:
                self._sample_nvml() if self.nvml is not None else self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.5)

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
        sm = [s[0] for s in self.samples]
        bits = 0
        for s in self.samples:
            bits |= s[2]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(s[1] for s in self.samples), "reasons": [n for n, b in self.REASONS if bits & b],
                "samples": len(self.samples), "gpus": self.indices, "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle)
def cpu_oracle_throughput(envs_per_core, control_steps, render, seed=0):
    """Times the CPU oracle (float32 physics build + the CPU raster oracle when `render`) on all host cores: one worker PROCESS per core,
    each stepping its share of the sample (the reference vectorises its CPU backend the same way, one env process per core,
    mani_skill/examples/benchmarking/gpu_sim.py:72-84).  Returns (env-steps/s, cores, seconds)."""
    from oracle import oracle as _o
    from oracle import raster as _r
    _o.build()
    _r.lib()
    cores = usable_cores()
    start_at = time.time() + 8.0 + 0.05 * envs_per_core + 0.05 * cores  # imports + world construction + warm-up happen before this instant
    env_ = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "oracle", "cpu_worker.py"), str(envs_per_core), str(control_steps), str(seed + i),
                               repr(start_at), "rgbd" if render else "state"], stdout=subprocess.PIPE, text=True, env=env_) for i in range(cores)]
    times = [float(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    dt = max(times)
    return envs_per_core * cores * control_steps / dt, cores, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot be installed here (sapien/PhysX absent), so the
    CPU restatement (oracle/: physics + rasteriser, kind="port") is timed on the host cores with every core it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    render = "rgb" in args.obs_mode
    per_core = 32 if render else 128
    t0 = time.time()
    v, cores, dt = cpu_oracle_throughput(per_core, max(args.steps, 1), render)
    sample = per_core * cores
    line = {
        "impl": "reference", "metric": f"env steps/sec (PickCube-v1, {args.obs_mode})", "value": v, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(n=args.num_envs, mode=args.obs_mode), "sample_envs": sample, "substeps_per_step": 5,
                   "renders": render, "setup_s": time.time() - t0 - dt},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} sub-scenes ({per_core} per core) x {args.steps} control steps (5 substeps each"
                                   f"{' + one 128x128 CPU render' if render else ''}) of the same workload, CPU oracle f32, {cores} worker processes"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def pin_rank_to_cores(local_rank, world_size):
    """Each rank gets its own slice of the host cores that are local to ITS GPU (NVML's CPU affinity of the device = the NUMA node its PCIe
    root hangs off), shared evenly with the other ranks whose GPUs sit on the same node: the pinned host buffers of the end-to-end loop are
    then first-touched on the node the GPU copies into.  Falls back to contiguous slices of the usable cores.  Returns (cores, how)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        return None, "unpinned"
    if world_size <= 1:
        return None, "unpinned"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        words = max(cores) // 64 + 1
        masks = []
        for g in range(world_size):
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(g).uuid))
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(g)
            aff = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            masks.append(frozenset(64 * w + b for w, word in enumerate(aff) for b in range(64) if (int(word) >> b) & 1) & frozenset(cores))
        mine = masks[local_rank]
        peers = [g for g in range(world_size) if masks[g] == mine]
        per = len(mine) // len(peers)
        if per >= 2:
            lst = sorted(mine)
            k = peers.index(local_rank)
            os.sched_setaffinity(0, lst[k * per:(k + 1) * per])
            return per, "nvml-numa"
    except Exception:
        pass
    try:
        per = len(cores) // world_size
        if per >= 2:
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
            return per, "contiguous"
    except Exception:
        pass
    return None, "unpinned"


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)   # before anything queries device properties: a rank never initialises another rank's GPU
    pinned_cores, pinned_how = pin_rank_to_cores(local_rank, world_size)
    torch.set_num_threads(max(1, min(4, pinned_cores or 4)))
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    import maniskill_b200 as ms
    from maniskill_b200.backend import BUF_ALL
    from maniskill_b200.dist import shard_seeds
    t_build0 = time.time()
    steps, warmup = args.steps, max(args.warmup, 3)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phases = {}

    def make(obs_mode, n, task="PickCube-v1"):
        t0 = time.time()
        env = ms.make(task, num_envs=n, obs_mode=obs_mode, device=dev)
        t1 = time.time()
        venv = ms.ManiSkillVectorEnv(env, auto_reset=not args.no_auto_reset)
        venv.reset(seed=shard_seeds(2022, n * world_size, rank, world_size))  # seeds keep the global env id
        torch.cuda.synchronize()
        phases.setdefault("make_s", []).append(round(t1 - t0, 2))
        phases.setdefault("reset_s", []).append(round(time.time() - t1, 2))
        return env, venv

    def flat_state(o):
        return o["state"] if isinstance(o, dict) else o

    def timed_value(env, venv, n, k, gather, sampler=None):
        """k control steps with device-resident actions, L2 flushed between steps, CUDA events around every step.  With `gather`, the
        all-gather of step i's flattened state runs asynchronously (NCCL's stream) under step i + 1 and is awaited before the buffer is
        reused (two buffers).  Returns (summed device ms, sorted per-step ms, launches, wall seconds)."""
        A = env.action_dim
        for _ in range(warmup):
            venv.step(2 * torch.rand((n, A), device=dev, generator=gen) - 1)
        actions_all = 2 * torch.rand((k, n, A), device=dev, generator=gen) - 1
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        gbuf, gwork, gt = None, [None, None], 0.0
        if gather:
            gbuf = [torch.empty((world_size * n, gather), dtype=torch.float32, device=dev) for _ in range(2)]
        w = env.scene.world
        barrier()
        l0 = w.kernel_launches
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            barrier()
            t0 = time.perf_counter()
            torch.cuda.nvtx.range_push("timed")  # ncu --nvtx --nvtx-include "timed/" captures exactly these launches
            for i in range(k):
                flush.fill_(float(i))  # evict L2 between timed iterations
                ev0[i].record()
                o, r, te, tr, info = venv.step(actions_all[i])
                if gather:
                    if gwork[i % 2] is not None:
                        gwork[i % 2].wait()
                    gwork[i % 2] = dist.all_gather_into_tensor(gbuf[i % 2], flat_state(o).contiguous(), async_op=True)
                ev1[i].record()
            for wk in gwork:
                if wk is not None:
                    wk.wait()
            torch.cuda.nvtx.range_pop()
            t_loop = time.perf_counter() - t0   # this rank's host time in the loop (before the closing barrier)
            barrier()
            t_wall = time.perf_counter() - t0
        ms_steps = sorted(a.elapsed_time(b) for a, b in zip(ev0, ev1))
        return sum(ms_steps), ms_steps, w.kernel_launches - l0, t_wall, t_loop

    def timed_e2e(env, venv, n, k):
        """The same k steps end to end through the public API with HOST buffers: actions from pinned host memory every step, the step's
        observation (state + images), reward and done flags read back to pinned host memory every step.  The device->host copies of step i
        run on a copy stream under step i + 1 (snapshot of the render targets taken on the compute stream); the host waits for step i - 1's
        results before it issues step i + 1, and for the last ones before the clock stops."""
        A = env.action_dim
        visual = env.obs_mode != "state"
        cpu_actions = 2 * torch.rand((k, n, A)) - 1
        h_act = [torch.empty((n, A), dtype=torch.float32).pin_memory() for _ in range(2)]
        host = []
        for _ in range(2):
            hb = dict(state=torch.empty((n, 2 * 9 + 24), dtype=torch.float32).pin_memory(), rew=torch.empty(n, dtype=torch.float32).pin_memory(),
                      done=torch.empty((2, n), dtype=torch.bool).pin_memory())
            if visual:
                hb.update(rgb=torch.empty((n, 128, 128, 3), dtype=torch.uint8).pin_memory(), depth=torch.empty((n, 128, 128, 1), dtype=torch.int16).pin_memory())
            host.append(hb)
        d2h = sum(t.numel() * t.element_size() for t in host[0].values())
        copy_stream = torch.cuda.Stream(device=dev)
        cev = [torch.cuda.Event() for _ in range(2)]
        barrier()
        t0 = time.perf_counter()
        for i in range(k):
            b = i % 2
            h_act[b].copy_(cpu_actions[i])
            a = h_act[b].to(dev, non_blocking=True)
            o, r, te, tr, info = venv.step(a)
            dn = torch.stack((te, tr))
            if visual:
                sd = o["sensor_data"]["base_camera"]
                parts = dict(state=o["state"], rew=r, done=dn, rgb=sd["rgb"].clone(), depth=sd["depth"].clone())  # snapshot: the next render overwrites the targets
            else:
                parts = dict(state=o, rew=r, done=dn)
            ready = torch.cuda.Event()
            ready.record()
            if i >= 1:
                cev[1 - b].synchronize()          # the host now holds step i-1's observation / reward / done
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                for name, t in parts.items():
                    host[b][name].copy_(t, non_blocking=True)
                    t.record_stream(copy_stream)
                cev[b].record(copy_stream)
        cev[(k - 1) % 2].synchronize()
        torch.cuda.synchronize()
        t_loop = time.perf_counter() - t0
        barrier()
        return time.perf_counter() - t0, t_loop, n * A * 4, d2h

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def check_overflow(env, what):
        reasons = env.scene.world.overflow_reasons()
        if reasons:
            raise SystemExit(f"bench.py: capacity overflow during {what}: {reasons} -- contacts or constraint rows were dropped, the run is invalid")
        return 0

    # ================================================================ headline: state + rgb + depth
    n_envs, mode = args.num_envs, args.obs_mode
    env, venv = make(mode, n_envs)
    visual = env.obs_mode != "state"
    build_s = time.time() - t_build0
    gather_dim = (2 * 9 + 24) if (world_size > 1 and not args.no_gather_obs) else 0
    sampler = ClockSampler(range(world_size)) if rank == 0 else None
    t_dev_ms, step_ms, launches, t_wall, t_loop = timed_value(env, venv, n_envs, steps, gather_dim, sampler)
    t_e2e, t_e2e_loop, h2d, d2h = timed_e2e(env, venv, n_envs, steps)
    check_overflow(env, "the headline run")
    # per-rank spread (host-side stragglers show up here): max and min over ranks of the loop times
    t_dev_ms_max, t_e2e_max, t_wall_max, t_loop_max, t_e2e_loop_max = reduce_max([t_dev_ms, t_e2e, t_wall, t_loop, t_e2e_loop])
    neg = reduce_max([-t_loop, -t_e2e_loop])
    total_envs = n_envs * world_size
    value = total_envs * steps / (t_dev_ms_max * 1e-3)
    e2e = total_envs * steps / t_e2e_max
    # ---------------- kernels alone (roofline numerators): CUDA events around back-to-back launches on the launching stream, no sync inside
    w = env.scene.world
    substeps = env._sim_steps_per_control
    k_iters = 20

    def time_launches(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k_iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k_iters

    step_ms_alone = time_launches(lambda: w.step(substeps, BUF_ALL))
    raster_ms = time_launches(lambda: env._sensors.capture()) if visual else None
    peak, peak_src = read_peaks()
    traffic = read_traffic()
    phys_bytes = n_envs * BYTES_PHYSICS_PER_ENV_SUBSTEP * substeps
    kernels = {"b2s_step": {"ms": step_ms_alone, "algorithmic_bytes_per_launch": phys_bytes, "achieved_GBps": phys_bytes / (step_ms_alone * 1e-3) / 1e9,
                            "frac": phys_bytes / (step_ms_alone * 1e-3) / 1e9 / peak, "traffic": traffic.get("b2s_step"),
                            "what": f"{substeps} substeps x (kin x2, collide, manifest, rowfill, solve) + fetch, one CUDA graph"}}
    if visual:
        px_bytes = n_envs * CAMERA_PIXELS * BYTES_PER_DELIVERED_PIXEL
        kernels["raster_kernel"] = {"ms": raster_ms, "algorithmic_bytes_per_launch": px_bytes, "achieved_GBps": px_bytes / (raster_ms * 1e-3) / 1e9,
                                    "frac": px_bytes / (raster_ms * 1e-3) / 1e9 / peak, "traffic": traffic.get("raster_kernel"),
                                    "what": "one CTA per (sub-scene, camera) image, writes rgb 3 B + depth 2 B per pixel"}
    dom = max(kernels, key=lambda k_: kernels[k_]["ms"])
    line = {
        "metric": f"env steps/sec (PickCube-v1, {mode})", "value": value, "unit": "env-steps/s", "n_gpus": world_size,
        "steps": steps, "warmup": warmup, "ms_per_step": t_dev_ms_max / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(n=n_envs, mode=mode), "num_envs_total": total_envs, "substeps_per_s": value * substeps,
                   "l2": "256 MiB write between timed steps (and every step writes more render-target bytes than L2 holds)",
                   "obs_all_gather": bool(gather_dim), "device_autoreset": bool(venv._device_autoreset), "overflow": 0,
                   "wall_s": t_wall_max, "step_ms_median_rank0": step_ms[len(step_ms) // 2], "step_ms_max_rank0": step_ms[-1],
                   "host_loop_s_max_over_ranks": t_loop_max, "host_loop_s_min_over_ranks": -neg[0],
                   "e2e_loop_s_max_over_ranks": t_e2e_loop_max, "e2e_loop_s_min_over_ranks": -neg[1],
                   "build_s": build_s, "cores_per_rank": pinned_cores, "core_pinning": pinned_how},
        "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "pinned host actions in; state + rgb + depth + reward + done out to pinned host memory every step, device->host copies of "
                       "step i overlapped with step i+1 (copy stream), host waits for step i-1 before issuing step i+1"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                     "traffic": kernels[dom]["traffic"], "peak_source": peak_src, "kernel": dom, "kernel_ms": kernels[dom]["ms"],
                     "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"],
                     "note": "kernel_ms = mean of 20 back-to-back launches between two CUDA events (no sync inside); traffic = dram bytes per launch "
                             "parsed from profiles/r02_traffic.json (ncu --set full), null when that file is absent"},
        "kernels": kernels,
    }
    if rank == 0 and sampler is not None:
        line["clocks"] = sampler.summary()
    env.close()
    del env, venv
    # ================================================================ side keys
    side = {}
    if not args.no_state_only and visual:
        env_s, venv_s = make("state", n_envs)
        ts, sms, ls, _, _ = timed_value(env_s, venv_s, n_envs, steps, gather_dim)
        te, _, h2d_s, d2h_s = timed_e2e(env_s, venv_s, n_envs, steps)
        check_overflow(env_s, "the state-only run")
        ts_m, te_m = reduce_max([ts, te])
        side["state_only"] = {"workload": f"PickCube-v1 num_envs={n_envs}/GPU obs_mode=state (BASELINE.json configs[1])",
                              "value": total_envs * steps / (ts_m * 1e-3), "ms_per_step": ts_m / steps, "e2e": total_envs * steps / te_m,
                              "h2d_bytes_per_step": h2d_s, "d2h_bytes_per_step": d2h_s, "gpu_launches": int(ls)}
        env_s.close()
        del env_s, venv_s
    if not args.no_scan:
        scan = {}
        for n in (1024, 16384):
            for m in ("state", mode) if visual else ("state",):
                k = 20 if m == "state" else 10
                env_n, venv_n = make(m, n)
                tn, _, _, _, _ = timed_value(env_n, venv_n, n, k, 0)
                check_overflow(env_n, f"the {n}-env scan")
                tn_m, = reduce_max([tn])
                scan[f"{n}:{m}"] = {"env_steps_per_s": n * world_size * k / (tn_m * 1e-3), "ms_per_step": tn_m / k, "steps": k}
                env_n.close()
                del env_n, venv_n
        side["num_envs_scan"] = scan
    if not args.no_other_configs:
        # BASELINE.json configs[2] and configs[3], device-timed like `value` (L2 flushed, auto-reset on); their task epilogues (evaluate /
        # reward / observation) are eager torch over the C-ABI buffers, not the fused kernel PickCube has
        other = {}
        cab_n = max(1, 2048 // world_size)
        for key, task, m, n in (("configs[2]", "PegInsertionSide-v1", "rgbd", 4096), ("configs[3]", "OpenCabinetDrawer-v1", "state", cab_n)):
            env_o, venv_o = make(m, n, task)
            to, _, lo, _, _ = timed_value(env_o, venv_o, n, 10, 0)
            check_overflow(env_o, f"the {task} run")
            to_m, = reduce_max([to])
            other[key] = {"workload": f"{task} num_envs={n}/GPU obs_mode={m}", "env_steps_per_s": n * world_size * 10 / (to_m * 1e-3),
                          "ms_per_step": to_m / 10, "steps": 10, "gpu_launches": int(lo)}
            env_o.close()
            del env_o, venv_o
        side["other_configs"] = other
    line["side"] = side
    if rank == 0:
        if not args.no_cpu_baseline and world_size == 1:
            tb = time.time()
            render = visual
            per_core = 16 if render else 64
            v, cores, dt = cpu_oracle_throughput(per_core, 100, render)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{per_core * cores} sub-scenes ({per_core} per core) x 100 control steps (5 substeps each"
                                              f"{' + one 128x128 CPU render' if render else ''}) of the same workload, CPU oracle f32, "
                                              f"{cores} worker processes (usable cores), {dt:.1f}s wall",
                                    "leg_s": time.time() - tb}
        else:
            line["cpu_baseline"] = None
        line["config"]["total_s"] = time.time() - t_build0
        line["config"]["build_phases_rank0"] = phases
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--num-envs", type=int, default=4096, help="sub-scenes per GPU")
    ap.add_argument("--obs-mode", default="state+rgb+depth", help="headline observation mode (BASELINE.json metric: state+RGBD); 'state' = configs[1]")
    ap.add_argument("--no-gather-obs", action="store_true", help="N > 1: skip the NCCL all-gather of the flattened state observation")
    ap.add_argument("--no-auto-reset", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-state-only", action="store_true", help="skip the state-only side measurement")
    ap.add_argument("--no-scan", action="store_true", help="skip the num_envs = 1024 / 16384 side measurements")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the PegInsertionSide-v1 rgbd / OpenCabinetDrawer-v1 side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
