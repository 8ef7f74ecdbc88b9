"""CPU: trajectory record / replay (maniskill_b200/trajectory.py) on the emulated backend -- file layout of the reference's
RecordEpisode (mani_skill/utils/wrappers/record.py:546-756), one episode per sub-scene flushed by the vector wrapper's partial resets,
and replays that reproduce the recorded final state (replay_trajectory.py:111-378: by seed + actions, and by stored env states)."""
import json

import numpy as np
import pytest
import torch

import maniskill_b200 as ms
from emu_world import EmuBackendWorld
from maniskill_b200.trajectory import RecordEpisode, load_trajectories, replay_trajectory


def test_record_layout_and_partial_flush(tmp_path):
    n = 3
    env = ms.make("PickCube-v1", num_envs=n, obs_mode="state", world_factory=EmuBackendWorld)
    env.max_episode_steps = 6
    rec = RecordEpisode(env, str(tmp_path), trajectory_name="demo", source_type="test", source_desc="random actions")
    venv = ms.ManiSkillVectorEnv(rec, auto_reset=True)
    venv.reset(seed=7)
    g = torch.Generator().manual_seed(0)
    for t in range(8):      # 6 steps -> truncation of all three, auto-reset, 2 more steps of the next episodes
        venv.step(2 * torch.rand(n, 8, generator=g) - 1)
    rec.close()
    meta, trajs = load_trajectories(str(tmp_path / "demo"))
    assert meta["env_info"]["env_id"] == "PickCube-v1" and meta["env_info"]["max_episode_steps"] == 6
    assert meta["env_info"]["env_kwargs"]["control_mode"] == "pd_joint_delta_pos" and meta["source_type"] == "test"
    eps = meta["episodes"]
    assert [e["episode_id"] for e in eps] == list(range(6)) and [e["elapsed_steps"] for e in eps] == [6, 6, 6, 2, 2, 2]
    assert all(e["control_mode"] == "pd_joint_delta_pos" and e["reset_kwargs"] == {} and e["success"] in (True, False) for e in eps)
    assert len({e["episode_seed"] for e in eps[:3]}) == 3      # per-sub-scene seeds derived from the reset seed (sapien_env.py:321)
    t0 = trajs["traj_0"]
    assert set(t0) == {"obs", "actions", "terminated", "truncated", "success", "env_states", "rewards"}
    assert t0["obs"].shape == (7, 42) and t0["actions"].shape == (6, 8) and t0["actions"].dtype == np.float32
    assert t0["terminated"].shape == (6,) and t0["truncated"].dtype == bool and t0["rewards"].shape == (6,) and t0["rewards"].dtype == np.float32
    assert t0["truncated"].tolist() == [False] * 5 + [True] and not trajs["traj_3"]["truncated"].any()     # TimeLimit at max_episode_steps
    assert set(t0["env_states"]) == {"actors", "articulations"} and t0["env_states"]["actors"]["cube"].shape == (7, 13)
    assert t0["env_states"]["articulations"]["panda"].shape == (7, 13 + 18)
    # the second episode of sub-scene 0 starts from the auto-reset state, not from where the first one ended
    t3 = trajs["traj_3"]
    assert t3["obs"].shape == (3, 42) and not np.allclose(t3["obs"][0], t0["obs"][-1])
    assert np.allclose(t3["env_states"]["articulations"]["panda"][0, 13 + 9:], 0)      # qvel is zero right after a reset
    raw = json.load(open(tmp_path / "demo.json"))
    assert raw["episodes"][0]["episode_seed"] == eps[0]["episode_seed"]


def test_empty_episodes_are_skipped_and_full_reset_flushes_everything(tmp_path):
    env = ms.make("PickCube-v1", num_envs=2, obs_mode="state", world_factory=EmuBackendWorld)
    rec = RecordEpisode(env, str(tmp_path), record_env_state=False, record_reward=False)
    rec.reset(seed=1)
    rec.reset(seed=2)                       # nothing happened since the first reset: no episode
    rec.step(torch.zeros(2, 8))
    rec.reset(options=dict(env_idx=torch.tensor([1])))       # flushes sub-scene 1 only
    rec.step(torch.zeros(2, 8))
    rec.reset(seed=3)                       # flushes both: sub-scene 0 with 2 steps, sub-scene 1 with 1
    rec.close()
    meta, trajs = load_trajectories(str(tmp_path / "trajectory"))
    assert [e["elapsed_steps"] for e in meta["episodes"]] == [1, 2, 1]
    assert set(trajs["traj_0"]) == {"obs", "actions", "terminated", "truncated", "success"}


@pytest.mark.parametrize("mode", ["seed_and_actions", "env_states", "first_env_state"])
def test_replay_reproduces_the_recording(tmp_path, mode):
    env = ms.make("PickCube-v1", num_envs=1, obs_mode="state", world_factory=EmuBackendWorld)
    rec = RecordEpisode(env, str(tmp_path))
    g = torch.Generator().manual_seed(5)
    for seed in (11, 12):
        rec.reset(seed=seed)
        for _ in range(5):
            rec.step(2 * torch.rand(1, 8, generator=g) - 1)
    rec.close()
    meta, trajs = load_trajectories(str(tmp_path / "trajectory"))
    assert [e["reset_kwargs"] for e in meta["episodes"]] == [{"seed": 11}, {"seed": 12}] and [e["episode_seed"] for e in meta["episodes"]] == [11, 12]
    env2 = ms.make("PickCube-v1", num_envs=1, obs_mode="state", world_factory=EmuBackendWorld)
    if mode != "seed_and_actions":
        env2.reset(seed=999)        # a different layout: only the stored states can bring the recorded one back
    res = replay_trajectory(env2, str(tmp_path / "trajectory"), use_env_states=mode == "env_states", use_first_env_state=mode == "first_env_state")
    assert len(res) == 2
    for r in res:
        assert r["final_state_error"] < 1e-5 and r["success"] == r["recorded_success"]
    with pytest.raises(ValueError):
        replay_trajectory(ms.make("PickCube-v1", num_envs=1, obs_mode="state", control_mode="pd_joint_pos", world_factory=EmuBackendWorld), str(tmp_path / "trajectory"))


def test_videos(tmp_path):
    """record.py:338-354,758-805: one frame before the first step and one after every step, sub-scenes tiled into int(sqrt(n)) rows, a video
    per flush (`max_steps_per_video`, `close`)."""
    cv2 = pytest.importorskip("cv2")
    env = ms.make("PickCube-v1", num_envs=4, obs_mode="state", render_mode="sensors", world_factory=EmuBackendWorld,
                  sensor_configs=dict(base_camera=dict(width=32, height=32)))
    rec = RecordEpisode(env, str(tmp_path), save_trajectory=False, save_video=True, video_fps=10, max_steps_per_video=3)
    rec.reset(seed=0)
    for _ in range(5):
        rec.step(torch.zeros(4, 8))
    rec.close()

    def frames(path):
        cap, out = cv2.VideoCapture(str(path)), []
        while True:
            ok, f = cap.read()
            if not ok:
                return out
            out.append(f)

    a, b = frames(tmp_path / "0.mp4"), frames(tmp_path / "1.mp4")
    assert len(a) == 4 and len(b) == 3           # s_0 + 3 steps, then s_3 + 2 steps flushed by close()
    assert a[0].shape == (2 * 32, 2 * 64, 3)     # 4 sub-scenes in 2 rows; each shows rgb + depth side by side
    assert a[0].std() > 5
    with pytest.raises(RuntimeError, match="render_mode"):
        RecordEpisode(ms.make("PickCube-v1", num_envs=1, obs_mode="state", world_factory=EmuBackendWorld), str(tmp_path), save_video=True)
