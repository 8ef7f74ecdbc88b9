"""CPU: the device per-env code compiled for the host (tests/emu) against the oracle -- proves the kernel logic
is the restated algorithm before any GPU time is spent. Both are float32 with contraction off => bit-level agreement."""
import numpy as np
import pytest

from scenarios import run_pick_cube


@pytest.mark.parametrize("emu_mode", [0, 1], ids=["fused", "pipelined"])
def test_emu_matches_oracle_f32_pick_cube(emu_mode):
    from maniskill_b200.scenes import pick_cube_scene
    cm = pick_cube_scene(6).compile()
    ref = run_pick_cube("oracle32", cm, 100)
    got = run_pick_cube("emu", cm, 100, emu_mode=emu_mode)
    assert np.abs(got["qpos"] - ref["qpos"]).max() < 1e-6
    assert np.abs(got["body"] - ref["body"]).max() < 1e-5
    rows = cm.link_rows["panda"]
    cube = cm.actor_rows["cube"]
    table = cm.actor_rows["table-workspace"]
    assert np.abs(got["world"].pair_impulse(cube, table) - ref["world"].pair_impulse(cube, table)).max() < 1e-6


@pytest.mark.parametrize("task,steps", [("PegInsertionSide-v1", 6), ("OpenCabinetDrawer-v1", 4)])
def test_emu_other_tasks_match_oracle(task, steps):
    """The pipelined substep on the scenes that use the other instantiations (per-env geometry overrides; two articulations,
    17 dofs, the 32-slot row records), device code compiled for the host against the oracle -- the CPU twin of
    tests/test_gpu_env.py::test_other_tasks_match_oracle."""
    from emu_world import EmuBackendWorld
    from scenarios import run_task_vs_oracle
    err_q, err_p, overflow = run_task_vs_oracle(task, steps, 3, world_factory=EmuBackendWorld)
    assert err_q < 1e-5 and err_p < 1e-5, (err_q, err_p)
    assert overflow == 0
