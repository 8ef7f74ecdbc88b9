"""TEST INFRASTRUCTURE ONLY: one CPU-baseline worker process (bench.py launches one per host core, the way the
reference vectorises its CPU backend with one process per env, mani_skill/examples/benchmarking/gpu_sim.py:72-84).
usage: cpu_worker.py <n_envs> <control_steps> <seed> [start_at] [state|rgbd]   -> prints seconds spent in the timed loop
With "rgbd" every control step also renders the task's 128x128 base camera of every sub-scene with the CPU raster oracle."""
import os
import sys
import time

sys.path[0] = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root instead of oracle/
import numpy as np  # noqa: E402

from maniskill_b200.scenes import PANDA_REST_QPOS, pick_cube_scene  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402

n_envs, steps, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cm = pick_cube_scene(n_envs).compile()
w = OracleWorld(cm, "f32")
rng = np.random.RandomState(seed)
q0 = PANDA_REST_QPOS + rng.normal(0, 0.02, (n_envs, 9))
q0[:, 7:] = 0.04
w.set_joint("qpos", q0)
w.set_joint("target_qpos", q0)


render = len(sys.argv) > 5 and sys.argv[5] == "rgbd"
if render:
    from maniskill_b200 import utils as U  # noqa: E402
    from maniskill_b200.render import build_visual_table, camera_desc  # noqa: E402
    from oracle import raster  # noqa: E402
    visuals = build_visual_table(cm, n_envs)
    cams = [camera_desc("base_camera", U.look_at([0.3, 0, 0.6], [-0.1, 0, 0.1]), 128, 128, np.pi / 2, 0.01, 100.0)]  # pick_cube.py:66-71


def observe():
    if render:
        raster.render(visuals, cams, w.rigid_body_data().astype(np.float32))


def act():
    tq = w.get_joint("qpos") + rng.uniform(-0.1, 0.1, (n_envs, 9))
    tq[:, 7:] = rng.uniform(-0.01, 0.04, (n_envs, 1))
    w.set_joint("target_qpos", tq)


act()
w.step(5)
observe()
# wait for the common start time so all workers overlap
start_at = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
while time.time() < start_at:
    time.sleep(0.001)
t0 = time.perf_counter()
for _ in range(steps):
    act()
    w.step(5)
    observe()
print(time.perf_counter() - t0, flush=True)
