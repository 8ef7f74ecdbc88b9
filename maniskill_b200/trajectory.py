"""Trajectory recording and replay: the on-disk format either side of the step path.

Mirror of `RecordEpisode` (mani_skill/utils/wrappers/record.py:113-826: trajectories and videos; no info overlay on the frames) and of the replay loop of
mani_skill/trajectory/replay_trajectory.py:111-378.  The reference writes `<name>.h5` + `<name>.json`:

    traj_<k>/obs                         [T+1, ...]  (dict observations become nested groups)
    traj_<k>/actions                     [T, A] float32
    traj_<k>/terminated, truncated       [T] bool        traj_<k>/success, fail  [T] bool (when the task reports them)
    traj_<k>/rewards                     [T] float32     (record_reward)
    traj_<k>/env_states/{actors,articulations}/<name>   [T+1, 13 | 13 + 2 dof]   (record_env_state)
    json: env_info {env_id, max_episode_steps, env_kwargs}, source_type, source_desc,
          episodes [{episode_id, episode_seed, control_mode, elapsed_steps, reset_kwargs, success, fail}]

The same keys and shapes are written here.  The container is HDF5 when `h5py` is importable; this image has no `h5py`, so the
fallback is a NumPy `.npz` archive whose member names are the HDF5 paths ("traj_0/env_states/actors/cube") -- `load_trajectories` reads
either into the same nested dicts.  One episode is flushed per sub-scene when it is reset (or on `close`), numbered in flush order.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional

import numpy as np
import torch

try:  # pragma: no cover - not available in this image
    import h5py
    if str(getattr(h5py, "__version__", "")).endswith("standin"):   # compat/site/h5py keeps whole files in memory: the npz path below streams
        h5py = None
except ImportError:
    h5py = None


def _to_numpy(x):
    if isinstance(x, dict):
        return {k: _to_numpy(v) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        x = x.detach()
        return x.cpu().numpy() if x.is_cuda else x.numpy().copy()   # never alias a caller-owned CPU buffer
    return np.array(x, copy=True)


def _index(x, idx):
    return {k: _index(v, idx) for k, v in x.items()} if isinstance(x, dict) else x[idx]


def _stack(frames: List):
    if isinstance(frames[0], dict):
        return {k: _stack([f[k] for f in frames]) for k in frames[0]}
    return np.stack(frames)


def _flatten(prefix: str, x, out: Dict[str, np.ndarray]):
    if isinstance(x, dict):
        for k, v in x.items():
            _flatten(f"{prefix}/{k}", v, out)
    else:
        out[prefix] = x


def _jsonable(x):
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, np.generic):
        return x.item()
    return x


class RecordEpisode:
    """record.py:113-327.  Wraps a BaseEnv; `reset` / `step` pass through and are recorded.  As in the reference the vector wrapper
    goes OUTSIDE (`ManiSkillVectorEnv(RecordEpisode(env, ...))`): its auto-reset then arrives here as a partial `reset(options=
    {"env_idx": ...})`, which flushes exactly the finished episodes.

    save_on_reset: flush the episodes of the sub-scenes being reset before resetting them (record.py:364-378).
    record_env_state: store `base_env.get_state_dict()` per frame so that a replay can restore states exactly."""

    def __init__(self, env, output_dir: str, save_trajectory: bool = True, trajectory_name: str = "trajectory", save_on_reset: bool = True,
                 record_reward: bool = True, record_env_state: bool = True, source_type: Optional[str] = None, source_desc: Optional[str] = None,
                 env_id: Optional[str] = None, env_kwargs: Optional[dict] = None, save_video: bool = False, video_fps: int = 30,
                 max_steps_per_video: Optional[int] = None):
        self.env = env
        self.base_env = getattr(env, "base_env", env)
        self.num_envs = self.base_env.num_envs
        self.output_dir = output_dir
        self.save_trajectory, self.save_on_reset = save_trajectory, save_on_reset
        self.record_reward, self.record_env_state = record_reward, record_env_state
        os.makedirs(output_dir, exist_ok=True)
        self._stem = os.path.join(output_dir, trajectory_name)
        self._episode_id = -1
        self._npz_started = False                      # npz fallback: datasets are appended to the zip at each flush
        self._h5 = h5py.File(self._stem + ".h5", "w") if (h5py is not None and save_trajectory) else None
        from .registration import REGISTERED_ENVS
        if env_id is None:
            env_id = next((k for k, (cls, _) in REGISTERED_ENVS.items() if cls is type(self.base_env)), type(self.base_env).__name__)
        kw = dict(obs_mode=self.base_env.obs_mode, reward_mode=self.base_env._reward_mode, control_mode=self.base_env.control_mode, num_envs=self.num_envs)
        kw.update(env_kwargs or {})
        self._json = dict(env_info=dict(env_id=env_id, max_episode_steps=self.base_env.max_episode_steps, env_kwargs=_jsonable(kw)),
                          source_type=source_type, source_desc=source_desc, episodes=[])
        self._frames: Optional[List[dict]] = None     # one dict of [num_envs, ...] arrays per frame since the oldest unflushed episode start
        self._start = np.zeros(self.num_envs, dtype=np.int64)   # record.py `env_episode_ptr`
        self._last_reset_kwargs: dict = {}
        # videos (record.py:338-354, 758-805): `env.render()` after every step (and once before the first), all sub-scenes tiled into
        # int(sqrt(num_envs)) rows, one `<k>.mp4` per flush (reset with save_on_reset, max_steps_per_video, close)
        self.save_video, self.video_fps, self.max_steps_per_video = save_video, video_fps, max_steps_per_video
        if save_video and self.base_env.render_mode is None:
            raise RuntimeError("save_video needs an env made with a render_mode")
        self.video_nrows = max(int(np.sqrt(self.num_envs)), 1)
        self.render_images: List[np.ndarray] = []
        self._video_id, self._video_steps = -1, 0

    # -------------------------------------------------------------- video
    def capture_image(self) -> np.ndarray:
        from .visualization import tile_images
        img = self.base_env.render()
        if img.dim() == 3:
            img = img[None]
        img = img[0] if len(img) == 1 else tile_images(list(img), nrows=self.video_nrows)
        return _to_numpy(img)

    def flush_video(self, name: Optional[str] = None, suffix: str = "", ignore_empty_transition: bool = True, save: bool = True):
        if not self.render_images or (ignore_empty_transition and len(self.render_images) == 1):
            return None
        path = None
        if save:
            self._video_id += 1
            video_name = name if name is not None else f"{self._video_id}" + (f"_{suffix}" if suffix else "")
            path = images_to_video(self.render_images, self.output_dir, video_name, fps=self.video_fps)
        self._video_steps = 0
        self.render_images = []
        return path

    # -------------------------------------------------------------- pass-through
    def __getattr__(self, name):
        return getattr(self.env, name)

    def _frame(self, obs, action, reward, terminated, truncated, info):
        f = dict(obs=_to_numpy(obs), action=_to_numpy(action), terminated=_to_numpy(terminated), truncated=_to_numpy(truncated))
        if self.record_reward:
            f["reward"] = _to_numpy(reward).astype(np.float32)
        if self.record_env_state:
            f["state"] = _to_numpy(self.base_env.get_state_dict())
        for k in ("success", "fail"):
            if k in info:
                f[k] = _to_numpy(info[k])
        return f

    def reset(self, seed=None, options: Optional[dict] = None, save: bool = True):
        idx = np.arange(self.num_envs) if not options or "env_idx" not in options else np.atleast_1d(_to_numpy(options["env_idx"])).astype(np.int64)
        if self.save_video and self.save_on_reset and self.num_envs == 1:     # record.py:364-366: with several sub-scenes videos run on
            self.flush_video(save=save)
        if self.save_trajectory and self.save_on_reset and self._frames is not None:
            self.flush_trajectory(env_idxs_to_flush=idx, save=save)
        obs, info = self.env.reset(seed=seed, options=options)
        if self.save_trajectory:
            N = self.num_envs
            first = self._frame(obs, np.zeros((N, self.base_env.action_dim), dtype=np.float32), np.zeros(N, dtype=np.float32),
                                np.ones(N, dtype=bool), np.ones(N, dtype=bool), dict(success=np.zeros(N, dtype=bool)))
            if self._frames is None:
                self._frames = [first]
            else:  # the frame a flushed episode ended on becomes the first frame of the next one (record.py:427-458)
                last = self._frames[-1]

                def replace(dst, src):
                    if isinstance(dst, dict):
                        for k in dst:
                            if k in src:
                                replace(dst[k], src[k])
                    else:
                        dst[idx] = src[idx]
                replace(last, first)
                self._start[idx] = len(self._frames) - 1
        self._last_reset_kwargs = dict(seed=_jsonable(seed)) if seed is not None else {}
        return obs, info

    def step(self, action):
        if self.save_video and self._video_steps == 0:
            self.render_images.append(self.capture_image())        # s_0, taken here so that repeated resets leave no empty videos
        obs, rew, terminated, truncated, info = self.env.step(action)
        if self.base_env.max_episode_steps is not None:
            # the env the reference's recorder wraps comes out of gym.make with the TimeLimit wrapper applied (registration.py:127-170)
            truncated = truncated | (self.base_env.elapsed_steps >= self.base_env.max_episode_steps)
        if self.save_trajectory:
            self._frames.append(self._frame(obs, action, rew, terminated, truncated, info))
        if self.save_video:
            self._video_steps += 1
            self.render_images.append(self.capture_image())
            if self.max_steps_per_video is not None and self._video_steps >= self.max_steps_per_video:
                self.flush_video()
        return obs, rew, terminated, truncated, info

    # -------------------------------------------------------------- record.py:546-756
    def flush_trajectory(self, env_idxs_to_flush=None, ignore_empty_transition: bool = True, save: bool = True):
        if self._frames is None:
            return
        idxs = np.arange(self.num_envs) if env_idxs_to_flush is None else np.asarray(env_idxs_to_flush)
        end = len(self._frames)
        flushed = []
        for e in idxs:
            start = int(self._start[e])
            if ignore_empty_transition and end - start <= 1:
                continue
            flushed.append(e)
            if not save:
                continue
            self._episode_id += 1
            fr = self._frames[start:end]
            data = dict(obs=_stack([_index(f["obs"], e) for f in fr]), actions=np.stack([f["action"][e] for f in fr[1:]]).astype(np.float32),
                        terminated=np.array([f["terminated"][e] for f in fr[1:]], dtype=bool), truncated=np.array([f["truncated"][e] for f in fr[1:]], dtype=bool))
            ep = dict(episode_id=self._episode_id, episode_seed=int(self.base_env._episode_seed[e]), control_mode=self.base_env.control_mode,
                      elapsed_steps=end - start - 1, reset_kwargs=dict(self._last_reset_kwargs) if self.num_envs == 1 else {})
            for k in ("success", "fail"):
                if all(k in f for f in fr[1:]):
                    data[k] = np.array([f[k][e] for f in fr[1:]], dtype=bool)
                    ep[k] = bool(data[k][-1])
            if self.record_env_state:
                data["env_states"] = _stack([_index(f["state"], e) for f in fr])
            if self.record_reward:
                data["rewards"] = np.array([f["reward"][e] for f in fr[1:]], dtype=np.float32)
            self._write(f"traj_{self._episode_id}", data)
            self._json["episodes"].append(_jsonable(ep))
        if flushed:
            self._start[flushed] = end - 1
            drop = int(self._start.min())
            if drop > 0:
                self._frames = self._frames[drop:]
                self._start -= drop
            if save:
                self._dump()

    def _write(self, name: str, data: dict):
        flat: Dict[str, np.ndarray] = {}
        _flatten(name, data, flat)
        if self._h5 is not None:  # pragma: no cover
            for path, arr in flat.items():
                big = path.rsplit("/", 1)[-1] in ("rgb", "depth", "seg")
                self._h5.create_dataset(path, data=arr, **(dict(compression="gzip", compression_opts=5) if big else {}))
        else:
            # `.npz` container (h5py missing): one zip member per dataset, appended at each flush and dropped from memory -- O(1) work and
            # host RAM per flushed episode, readable with np.load like a file written by np.savez_compressed
            import zipfile
            with zipfile.ZipFile(self._stem + ".npz", "a" if self._npz_started else "w", zipfile.ZIP_DEFLATED, allowZip64=True) as zf:
                for path, arr in flat.items():
                    with zf.open(path + ".npy", "w", force_zip64=True) as member:
                        np.lib.format.write_array(member, np.asanyarray(arr), allow_pickle=False)
            self._npz_started = True

    def _dump(self):
        with open(self._stem + ".json", "w") as f:
            json.dump(self._json, f, indent=2)
        if self._h5 is not None:  # pragma: no cover
            self._h5.flush()
        elif not self._npz_started:   # nothing flushed yet: still leave a (valid, empty) container next to the json
            np.savez_compressed(self._stem + ".npz")
            self._npz_started = True

    def close(self):
        if self.save_video and self.save_on_reset:
            self.flush_video()
        if self.save_trajectory:
            self.flush_trajectory()
            self._dump()
            if self._h5 is not None:  # pragma: no cover
                self._h5.close()
        if hasattr(self.env, "close"):
            self.env.close()


def images_to_video(images: List[np.ndarray], output_dir: str, video_name: str, fps: int = 30) -> str:
    """mani_skill/utils/visualization/misc.py `images_to_video` (imageio + ffmpeg there; OpenCV's mp4 writer here): [H, W, 3] uint8 RGB
    frames -> `<output_dir>/<video_name>.mp4`."""
    import cv2
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, video_name.replace(" ", "_").replace("\n", "_") + ".mp4")
    h, w = images[0].shape[:2]
    writer = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
    if not writer.isOpened():
        raise RuntimeError(f"cannot open a video writer for {path}")
    for im in images:
        writer.write(cv2.cvtColor(np.ascontiguousarray(im), cv2.COLOR_RGB2BGR))
    writer.release()
    return path


# ------------------------------------------------------------------------------------------------ reading + replay
def load_trajectories(path_stem: str):
    """-> (json dict, {"traj_<k>": nested dict of arrays}) from `<stem>.json` + `<stem>.h5` | `<stem>.npz`."""
    with open(path_stem + ".json") as f:
        meta = json.load(f)
    trajs: Dict[str, dict] = {}

    def put(path, arr):
        node = trajs
        parts = path.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = arr

    import zipfile
    h5 = path_stem + ".h5"
    if os.path.exists(h5) and not zipfile.is_zipfile(h5):  # pragma: no cover  (real HDF5)
        if h5py is None:
            raise RuntimeError("reading HDF5 trajectories needs h5py")
        with h5py.File(h5, "r") as f:
            f.visititems(lambda name, obj: put(name, obj[()]) if isinstance(obj, h5py.Dataset) else None)
    else:
        # `.npz`, or an `.h5` written through compat/site/h5py (the reference's RecordEpisode on the shim): both are zip archives of .npy members
        with np.load(h5 if os.path.exists(h5) else path_stem + ".npz") as z:
            for k in z.files:
                if k != "__attrs__.json":
                    put(k, z[k])
    return meta, trajs


def replay_trajectory(env, path_stem: str, use_env_states: bool = False, use_first_env_state: bool = False, count: Optional[int] = None,
                      allow_failure: bool = True):
    """replay_trajectory.py:111-378 for one sub-scene per episode: reset with the stored seed, optionally restore the stored first
    state, feed the stored actions (restoring the stored state after every step with `use_env_states`), and report per episode
    whether the replay ended in the recorded success flag and how far the final state is from the recorded one.

    env: a BaseEnv with num_envs == 1 and the control mode of the recording."""
    base = getattr(env, "base_env", env)
    if base.num_envs != 1:
        raise ValueError("replay one episode at a time: make the env with num_envs=1")
    meta, trajs = load_trajectories(path_stem)
    results = []
    dev = base.device
    to_t = lambda d: {k: to_t(v) for k, v in d.items()} if isinstance(d, dict) else torch.as_tensor(d, device=dev)[None]
    for ep in meta["episodes"][:count]:
        tr = trajs[f"traj_{ep['episode_id']}"]
        if ep["control_mode"] != base.control_mode:
            raise ValueError(f"episode {ep['episode_id']} was recorded with control mode {ep['control_mode']}, the env uses {base.control_mode}")
        env.reset(seed=ep["episode_seed"])
        states = tr.get("env_states")
        if (use_env_states or use_first_env_state):
            if states is None:
                raise ValueError("the trajectory holds no env_states")
            base.set_state_dict(to_t(_index(states, 0)))
        info = {}
        for t, a in enumerate(tr["actions"]):
            _, _, _, _, info = env.step(torch.as_tensor(a, device=dev)[None])
            if use_env_states:
                base.set_state_dict(to_t(_index(states, t + 1)))
        success = bool(_to_numpy(info["success"]).reshape(-1)[0]) if "success" in info else None
        err = None
        if states is not None:
            flat_now, flat_rec = {}, {}
            _flatten("", _to_numpy(base.get_state_dict()), flat_now)
            _flatten("", _index(states, len(tr["actions"])), flat_rec)
            err = max(float(np.abs(flat_now[k].reshape(-1) - flat_rec[k].reshape(-1)).max()) for k in flat_now)
        res = dict(episode_id=ep["episode_id"], success=success, recorded_success=ep.get("success"), final_state_error=err)
        if not allow_failure and ep.get("success") and not success:
            raise RuntimeError(f"replay of episode {ep['episode_id']} did not reach the recorded success")
        results.append(res)
    return results
