"""The reference's CPU simulation path (`sim_backend="physx_cpu"`, BASELINE.json configs[0]) run through the UNMODIFIED reference python with the CPU ORACLE
standing where SAPIEN's CPU PhysX stands (tests/cpu_sim_double.py) -- SURVEY.md section 8(c): "the reference's own semantic tests re-run against the oracle
through unchanged ManiSkill code".  Next to it the reference's GPU path runs on the emulated device code (tests/emu), so the reference's OWN CPU-vs-GPU tests
(tests/test_ik_controller.py, atol 5e-4) and a 100-substep rollout compare oracle and device code through the reference's two code paths (per-object getters /
setters + `px.get_contacts()` on one side, the `cuda_*` buffers + contact-impulse queries on the other).  The product itself has no CPU path: outside of
`installed()` the shim's `PhysxCpuSystem()` raises.  Skipped where /root/reference does not exist (the GPU box)."""
import os

import pytest
import torch

from test_reference_unmodified import REF, _run_reference_test, reference  # noqa: F401  (the fixture installs the shim and imports the reference)

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mani_skill")), reason="needs the reference checkout at /root/reference")


@pytest.fixture()
def cpu_backend(reference):  # noqa: F811
    from cpu_sim_double import installed
    with installed("f32"):
        yield reference


def test_float64_cpu_backend_against_the_float32_gpu_path(reference):  # noqa: F811
    """the same rollout with the FLOAT64 build of the oracle behind the CPU path: the float32 device code stays within north_star's 1e-4 of a double-precision
    restatement over 100 substeps of random actions (PickCube-v1)"""
    from cpu_sim_double import installed
    gym = reference
    with installed("f64"):
        cpu = gym.make("PickCube-v1", num_envs=1, obs_mode="state", sim_backend="physx_cpu")
        gpu = gym.make("PickCube-v1", num_envs=2, obs_mode="state", sim_backend="physx_cuda")
        cpu.reset(seed=4)
        gpu.reset(seed=4)
        one = lambda d, i: {kk: one(v, i) for kk, v in d.items()} if isinstance(d, dict) else d[i:i + 1]
        cpu.unwrapped.set_state_dict(one(gpu.unwrapped.get_state_dict(), 1))
        g = torch.Generator().manual_seed(1)
        worst = 0.0
        for _ in range(20):
            a = 2 * torch.rand((2, 8), generator=g) - 1
            og = gpu.step(a)[0]
            oc = cpu.step(a[1])[0]
            worst = max(worst, float((oc[0] - og[1]).abs().max()))
        assert worst < 1e-4, worst
        cpu.close()
        gpu.close()


def test_the_product_has_no_cpu_system(reference):  # noqa: F811
    from sapien import physx
    with pytest.raises(RuntimeError, match="no CPU simulation"):
        physx.PhysxCpuSystem()


@pytest.mark.parametrize("task", ["PickCube-v1", "PegInsertionSide-v1", "OpenCabinetDrawer-v1", "StackCube-v1", "PushCube-v1"])
def test_cpu_rollout_equals_the_gpu_path(cpu_backend, task):
    """BASELINE.json configs[0] (PickCube-v1, num_envs=1, CPU simulation, state observations) -- and the same for the other two tasks north_star names and two more --
    against a sub-scene of the same task on the GPU path, both through the reference's own env code, from the same state and with the same actions: observations
    (q, qdot, tcp / object / goal poses) and rewards within 1e-4 after 20 control steps = 100 substeps (north_star's tolerance; measured 1e-6 ... 4e-5).
    `reconfiguration_freq=0`: the reference rebuilds a single-env CPU scene on every reset by default (sapien_env.py), which would draw new peg / cabinet geometry."""
    gym = cpu_backend
    k = 2 if task == "PickCube-v1" else 0     # per-sub-scene geometry (peg, cabinet id) is drawn from seed 2022 + i: sub-scene 0 is the one a single env builds
    cpu = gym.make(task, num_envs=1, obs_mode="state", sim_backend="physx_cpu", reconfiguration_freq=0)
    gpu = gym.make(task, num_envs=4, obs_mode="state", sim_backend="physx_cuda")
    assert not cpu.unwrapped.gpu_sim_enabled and gpu.unwrapped.gpu_sim_enabled
    cpu.reset(seed=3)
    gpu.reset(seed=3)
    one = lambda d, i: {kk: one(v, i) for kk, v in d.items()} if isinstance(d, dict) else d[i:i + 1]
    cpu.unwrapped.set_state_dict(one(gpu.unwrapped.get_state_dict(), k))
    assert float((cpu.unwrapped.get_obs()[0] - gpu.unwrapped.get_obs()[k]).abs().max()) < 1e-6
    g = torch.Generator().manual_seed(0)
    A = cpu.action_space.shape[-1]
    for i in range(20):
        a = 2 * torch.rand((4, A), generator=g) - 1
        og, rg, _, _, ig = gpu.step(a)
        oc, rc, _, _, ic = cpu.step(a[k])
        assert float((oc[0] - og[k]).abs().max()) < 1e-4 and float((rc[0] - rg[k]).abs()) < 1e-4, i
        if "is_grasped" in ic:
            assert bool(ic["is_grasped"][0]) == bool(ig["is_grasped"][k])
    cpu.close()
    gpu.close()


def test_grasp_detection_agrees_between_the_contact_apis(cpu_backend):
    """`agent.is_grasping` reads `px.get_contacts()` on the CPU path (sapien_utils.py:216-262) and the contact-pair impulse queries on the GPU path
    (scene.py:741-801): the cube placed between the open fingers, gripper closing -- both report the grasp at the same step."""
    gym = cpu_backend
    cpu = gym.make("PickCube-v1", num_envs=1, obs_mode="state", sim_backend="physx_cpu", robot_init_qpos_noise=0.0)
    gpu = gym.make("PickCube-v1", num_envs=2, obs_mode="state", sim_backend="physx_cuda", robot_init_qpos_noise=0.0)
    cpu.reset(seed=1)
    gpu.reset(seed=1)
    sd = gpu.unwrapped.get_state_dict()
    tcp = gpu.unwrapped.agent.tcp.pose.p
    sd["actors"]["cube"][:, :3] = tcp
    sd["actors"]["cube"][:, 3:7] = torch.tensor([1.0, 0, 0, 0])
    sd["actors"]["cube"][:, 7:] = 0
    gpu.unwrapped.set_state_dict(sd)
    one = lambda d, i: {k: one(v, i) for k, v in d.items()} if isinstance(d, dict) else d[i:i + 1]
    cpu.unwrapped.set_state_dict(one(sd, 0))
    close = torch.zeros(8)
    close[7] = -1.0
    seen = []
    for i in range(12):
        _, _, _, _, ig = gpu.step(close.expand(2, -1))
        _, _, _, _, ic = cpu.step(close)
        assert bool(ic["is_grasped"][0]) == bool(ig["is_grasped"][0]), i
        seen.append(bool(ic["is_grasped"][0]))
    assert not seen[0] and sum(seen) >= 5, [int(s) for s in seen]
    cpu.close()
    gpu.close()


@pytest.mark.parametrize("fn,args", [
    ("test_ik_controller.py:test_pd_ee_delta_controller", ("pd_ee_delta_pose",)), ("test_ik_controller.py:test_pd_ee_delta_controller", ("pd_ee_target_delta_pose",)),
    ("test_ik_controller.py:test_pd_ee_delta_controller", ("pd_ee_delta_pos",)), ("test_ik_controller.py:test_pd_ee_delta_controller", ("pd_ee_target_delta_pos",)),
    ("test_ik_controller.py:test_pd_ee_controller", ("pd_ee_pose",)),
    ("test_envs.py:test_env_seeded_sequence_reset", ()), ("test_envs.py:test_states", ("PickCube-v1",)), ("test_envs.py:test_env_control_modes", ("PickCube-v1", "pd_joint_delta_pos")),
    ("test_envs.py:test_env_control_modes", ("PickCube-v1", "pd_ee_delta_pose")), ("test_envs.py:test_robots", ("PickCube-v1", "panda")),
    ("test_envs.py:test_env_raise_value_error_for_nan_actions", ()), ("test_envs.py:test_envs_obs_modes", ("PickCube-v1", "state_dict")),
    ("test_envs.py:test_envs_obs_modes", ("PickCube-v1", "state")), ("test_envs.py:test_states", ("StackCube-v1",)), ("test_envs.py:test_states", ("PegInsertionSide-v1",)),
    ("test_examples.py:test_demo_random_action", ()),
], ids=lambda v: "-".join(v) if isinstance(v, tuple) else str(v))
def test_reference_own_cpu_tests(cpu_backend, fn, args):
    """/root/reference/tests executed as they are.  tests/test_ik_controller.py is the reference's one numeric CPU-vs-GPU test (end-effector pose after 5 / 20
    control steps of the pd_ee_* controllers, atol 5e-4): here oracle vs emulated device code.  tests/test_envs.py: same-seed determinism of a reset / step
    sequence, state get / set round trip (atol 1e-4), control modes, robots, NaN actions, observation modes -- on the CPU path."""
    module_file, _, fn = fn.rpartition(":")
    _run_reference_test(module_file, fn, *args)
