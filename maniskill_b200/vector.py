"""ManiSkillVectorEnv -- mirror of mani_skill/vector/wrappers/gymnasium.py:18-199 together with the TimeLimit
truncation the registry adds (mani_skill/utils/registration.py:127-170): auto-reset of finished sub-scenes via the
partial reset ``reset(options={"env_idx": ...})``, ``final_info`` / ``final_observation`` bookkeeping."""
from __future__ import annotations

from typing import Optional

import torch


def _clone_tree(x):
    """Deep clone of a tensor / nested dict of tensors (mani_skill/utils/common.py `torch_clone_dict`)."""
    if isinstance(x, dict):
        return {k: _clone_tree(v) for k, v in x.items()}
    return x.clone() if isinstance(x, torch.Tensor) else x


class ManiSkillVectorEnv:
    def __init__(self, env, auto_reset: bool = True, ignore_terminations: bool = False, max_episode_steps: Optional[int] = None,
                 record_metrics: bool = False, device_autoreset: Optional[bool] = None):
        self._env = env
        self.num_envs = env.num_envs
        self.auto_reset = auto_reset
        self.ignore_terminations = ignore_terminations
        self.max_episode_steps = max_episode_steps if max_episode_steps is not None else env.max_episode_steps
        self.device = env.device
        # gymnasium.py:79-88: running episode statistics reported as infos["episode"]
        self.record_metrics = record_metrics
        if record_metrics:
            self.success_once = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
            self.fail_once = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
            self.returns = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        # Device-side auto-reset (b2s_pick_task_autoreset) when the env offers it for this configuration; `device_autoreset=False` keeps
        # the reference's python flow (gymnasium.py:160-176), which is also what every other task uses.
        self._device_autoreset = bool(device_autoreset if device_autoreset is not None else True) and auto_reset and not record_metrics \
            and hasattr(env, "supports_device_autoreset") and env.supports_device_autoreset() \
            and self.max_episode_steps is not None and env.enable_time_limit(self.max_episode_steps)

    @property
    def base_env(self):
        return self._env

    @property
    def unwrapped(self):
        return self._env

    def reset(self, *, seed=None, options=None):
        obs, info = self._env.reset(seed=seed, options=dict() if options is None else options)
        if self.record_metrics:  # gymnasium.py:108-124: statistics restart for the sub-scenes being reset
            idx = options["env_idx"] if options is not None and "env_idx" in options else slice(None)
            self.success_once[idx] = False
            self.fail_once[idx] = False
            self.returns[idx] = 0
        return obs, info

    def _update_metrics(self, rew, infos):
        """Running statistics of the current episodes (gymnasium.py:131-145): undiscounted return, whether success / failure has
        been seen so far, length, mean reward."""
        self.returns += rew
        stats = {}
        for flag, acc in (("success", "success_once"), ("fail", "fail_once")):
            if flag in infos:
                setattr(self, acc, getattr(self, acc) | infos[flag])
                stats[acc] = getattr(self, acc).clone()
        stats["return"] = self.returns.clone()
        stats["episode_len"] = self._env.elapsed_steps.clone()
        stats["reward"] = stats["return"] / stats["episode_len"]
        return stats

    def step(self, actions):
        if self._device_autoreset:
            # the whole of what follows (time limit, dones, final_info / final_observation, partial reset) on the device, no host sync
            return self._env.step_autoreset(actions, ignore_terminations=self.ignore_terminations)
        obs, rew, terminations, truncations, infos = self._env.step(actions)
        if self.max_episode_steps is not None:
            # TimeLimitWrapper.step (registration.py:160-168)
            truncations = self._env.elapsed_steps >= self.max_episode_steps
        else:
            truncations = torch.zeros_like(terminations)
        stats = self._update_metrics(rew, infos) if self.record_metrics else None
        if self.ignore_terminations:
            terminations = torch.zeros_like(terminations)
            if stats is not None:  # what the episode looked like at its last step (gymnasium.py:149-156)
                stats.update({f"{k}_at_end": infos[k].clone() for k in ("success", "fail") if k in infos})
        if stats is not None:
            infos["episode"] = stats
        dones = torch.logical_or(terminations, truncations)
        if dones.any() and self.auto_reset:
            final_obs = _clone_tree(obs)   # gymnasium.py:165 `torch_clone_dict(obs)`: image entries alias the render targets the reset re-renders
            env_idx = torch.arange(0, self.num_envs, device=self.device)[dones]
            final_info = _clone_tree(infos)
            obs, infos = self.reset(options=dict(env_idx=env_idx))
            infos["final_info"] = final_info
            infos["_final_info"] = dones
            infos["final_observation"] = final_obs
            infos["_final_observation"] = dones
            infos["_elapsed_steps"] = dones
        return obs, rew, terminations, truncations, infos

    def close(self):
        self._env.close()
