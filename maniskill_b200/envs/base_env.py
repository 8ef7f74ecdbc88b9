"""BaseEnv -- host-side mirror of mani_skill/envs/sapien_env.py:45 (``BaseEnv``) and mani_skill/envs/scene.py:40
(``ManiSkillScene``) for the hot path: same public method names, argument meaning and return structure
(``reset(seed, options) -> (obs, info)``, ``step(action) -> (obs, reward, terminated, truncated, info)``), same
seeding scheme (main RNG 2022+i, per-episode RNG, ``torch.random.fork_rng`` + ``manual_seed(episode_seed[0])``,
sapien_env.py:321,857-1021), same partial-reset semantics through ``scene._reset_mask`` and the same 5-substep
control loop (sapien_env.py:1073-1132) -- executed by ONE fused kernel launch instead of 5 ``px.step()`` calls.

The backend is ``maniskill_b200.backend.World`` (CUDA, no CPU path).  Tests may inject another object with the same
interface through ``world_factory`` (tests/ only).
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch

from .. import utils as U
from ..backend import (BUF_ALL, BUF_APPLY_ALL, BUF_QF, BUF_QPOS, BUF_QVEL, BUF_RIGID, BUF_ROOT_POSE, BUF_TARGET_QPOS,
                       BUF_TARGET_QVEL)
from ..model import CompiledModel, SceneDesc, SimParams
from ..observations import parse_obs_mode, sensor_data_to_pointcloud
from ..building import Device
from ..visualization import camera_observations_to_images, tile_images
from ..structs import Actor, Articulation, Pose


class Scene:
    """What ``env.scene`` is in the reference (ManiSkillScene): owns the physics world + named actors/articulations."""
    BUF_RIGID, BUF_ROOT_POSE, BUF_QPOS, BUF_QVEL, BUF_QF, BUF_TARGET_QPOS, BUF_TARGET_QVEL = (
        BUF_RIGID, BUF_ROOT_POSE, BUF_QPOS, BUF_QVEL, BUF_QF, BUF_TARGET_QPOS, BUF_TARGET_QVEL)

    def __init__(self, world, cm: CompiledModel, desc: SceneDesc):
        self.world = world
        self._px = None
        self.cm = cm
        self.num_envs = world.n_envs
        self.device = world.device
        self.timestep = cm.scalars["dt"]
        self._reset_mask = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        self._reset_all = True  # python-side knowledge that the mask is all-true: setters then skip the boolean gather (a device sync)
        self._dirty = 0
        self.actors: Dict[str, Actor] = {}
        self.articulations: Dict[str, Articulation] = {}
        n_link = cm.scalars["n_link"]
        for act in desc.actors:
            if act.body_type == "static":
                continue
            a = Actor(self, act.name, cm.actor_rows[act.name], act.body_type, cm.actor_fb[act.name],
                      Pose.create(act.initial_pose, self.device))
            a.hidden = act.hidden
            self.actors[act.name] = a
        for art in desc.articulations:
            ai = cm.art_index[art.name]
            d0, d1 = cm.art_dof_start[ai], cm.art_dof_start[ai + 1]
            qlim = cm.arrays["dof_limit"].reshape(-1, 2)[d0:d1]
            self.articulations[art.name] = Articulation(self, art.name, ai, cm.link_rows[art.name], cm.dof_names[art.name], qlim)
        self._query_cache = {}

    @property
    def _reset_mask(self):
        """Which sub-scenes the setters act on (mani_skill/envs/scene.py `_reset_mask`).  Assigning a new mask drops the python-side
        "all sub-scenes" shortcut; code that knows the mask is all-true sets `_reset_all = True` afterwards."""
        return self.__dict__["_reset_mask_tensor"]

    @_reset_mask.setter
    def _reset_mask(self, mask):
        self.__dict__["_reset_mask_tensor"] = mask
        self._reset_all = False

    @property
    def px(self):
        """`scene.px` of the reference (scene.py:61-63): the `PhysxGpuSystem`-shaped view of the world, built on first use."""
        if self._px is None:
            from ..physx_shim import PhysxGpuSystem
            cm = self.cm
            names = [""] * self.world.n_rows
            for art, rows in cm.link_rows.items():
                for link, r in rows.items():
                    names[r] = f"{art}_{link}"
            for act, r in cm.actor_rows.items():
                if r >= 0:
                    names[r] = act
            self._px = PhysxGpuSystem(self.world, names, [(a, len(cm.dof_names[a])) for a in cm.art_index])
        return self._px

    # ---- mani_skill/envs/scene.py:379-380
    def step(self, substeps=1, fetch_mask=0):
        self.world.step(substeps, fetch_mask)

    # ---- mani_skill/envs/scene.py:950-986
    def _gpu_apply_all(self):
        if self._dirty:
            self.world.apply(self._dirty)
            self._dirty = 0

    def _gpu_fetch_all(self):
        self.world.fetch(BUF_ALL)

    # ---- mani_skill/envs/scene.py:741-801
    def get_pairwise_contact_impulses(self, obj1, obj2):
        key = (obj1.row, obj2.row)
        if key not in self._query_cache:
            self._query_cache[key] = self.world.create_contact_query([key])
        return self.world.query_contact_impulses(self._query_cache[key])[:, 0]

    def get_pairwise_contact_forces(self, obj1, obj2):
        return self.get_pairwise_contact_impulses(obj1, obj2) / self.timestep

    # ---- mani_skill/utils/structs/base.py:116-136 (px.gpu_create_contact_body_impulse_query / gpu_query_contact_body_impulses)
    def get_net_contact_impulses(self, obj):
        key = (obj.row, -2)  # backend.ANY_BODY
        if key not in self._query_cache:
            self._query_cache[key] = self.world.create_contact_query([key])
        return self.world.query_contact_impulses(self._query_cache[key])[:, 0]

    def get_net_contact_forces(self, obj):
        return self.get_net_contact_impulses(obj) / self.timestep

    def get_sim_state(self):
        state = {"actors": {}, "articulations": {}}
        for k, a in self.actors.items():
            if a.px_body_type == "static":
                continue
            state["actors"][k] = a.get_state()
        for k, a in self.articulations.items():
            state["articulations"][k] = a.get_state()
        return state

    def set_sim_state(self, state, env_idx=None):
        for k, s in state["actors"].items():
            self.actors[k].set_state(s, env_idx)
        for k, s in state["articulations"].items():
            self.articulations[k].set_state(s, env_idx)


class _GraphedEpilogue:
    """`BaseEnv._epilogue` captured into a CUDA graph after a few eager runs (which also create the contact queries and cached
    constants the task code allocates lazily).  The captured tensors live in the graph's memory pool; every replay overwrites them, so
    callers get clones (the eager path returns fresh tensors too)."""
    WARMUP = 2

    def __init__(self, env):
        self.env, self.graph, self.out, self.action, self.runs = env, None, None, None, 0

    def run(self, action):
        env = self.env
        if action is None:   # `step(None)` (no new action): nothing to feed the captured action buffer with -- eager
            return env._epilogue(None)
        if self.graph is None:
            if self.runs < self.WARMUP:
                self.runs += 1
                return env._epilogue(action)
            self.action = action.clone()
            torch.cuda.synchronize(env.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.out = env._epilogue(self.action)
            self.graph = graph
        self.action.copy_(action)
        self.graph.replay()
        return _clone_tree(self.out)


def _clone_tree(x):
    if isinstance(x, dict):
        return {k: _clone_tree(v) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return type(x)(_clone_tree(v) for v in x)
    return x.clone() if isinstance(x, torch.Tensor) else x


class BaseEnv:
    """Mirror of mani_skill.envs.sapien_env.BaseEnv for GPU simulation (num_envs >= 1, sim_backend physx_cuda)."""

    # sapien_env.py:124: the named modes + "any_textures" = every '+'-combination of the textures the shader writes (and state flags)
    SUPPORTED_OBS_MODES = ("state", "state_dict", "none", "sensor_data", "any_textures", "pointcloud")
    SUPPORTED_REWARD_MODES = ("normalized_dense", "dense", "sparse", "none")
    SUPPORTED_RENDER_MODES = ("rgb_array", "sensors", "all")   # "human" needs the interactive viewer, which is out of scope
    max_episode_steps: Optional[int] = None

    def __init__(self, num_envs: int = 1, obs_mode: Optional[str] = None, reward_mode: Optional[str] = None,
                 control_mode: Optional[str] = None, sim_config: Optional[dict] = None, device: Union[str, torch.device, None] = None,
                 world_factory=None, sensor_configs: Optional[dict] = None, enable_cameras: Optional[bool] = None, fused: bool = True,
                 enhanced_determinism: bool = False, reconfiguration_freq: Optional[int] = None, render_mode: Optional[str] = None,
                 sensor_outputs: str = "auto"):
        self.num_envs = num_envs
        if sensor_outputs not in ("auto", "raw", "compact"):
            raise ValueError("sensor_outputs must be 'auto', 'raw' or 'compact'")
        self._sensor_outputs = sensor_outputs
        self._obs_mode = "state" if obs_mode is None else obs_mode
        self.obs_mode_struct = parse_obs_mode(self._obs_mode)   # raises NotImplementedError for unknown / unsupported textures
        self._reward_mode = self.SUPPORTED_REWARD_MODES[0] if reward_mode is None else reward_mode   # sapien_env.py:300-304
        if self._reward_mode not in self.SUPPORTED_REWARD_MODES:
            raise NotImplementedError(f"Unsupported reward mode: {self._reward_mode}")
        self.sim_params = SimParams.from_config(sim_config)
        self._sim_freq, self._control_freq = self.sim_params.sim_freq, self.sim_params.control_freq
        if self._sim_freq % self._control_freq != 0:
            raise ValueError("sim_freq must be divisible by control_freq")  # sapien_env.py:279-283 warns; we are strict
        self._sim_steps_per_control = self._sim_freq // self._control_freq
        self._world_factory = world_factory
        self._state_version = 0
        self._epilogue_runner = None
        self._requested_device = device
        self._main_seed = None
        self._episode_seed = np.zeros(num_envs, dtype=np.int64)
        self._batched_main_rng = None
        self._batched_episode_rng = None
        self._episode_rng = None
        # sapien_env.py:115-118: with it, resets without a seed re-seed the episode RNG of the sub-scenes being reset from their main RNG
        self._enhanced_determinism = bool(enhanced_determinism)
        self._sensor_overrides = sensor_configs or {}
        self._visual = self.obs_mode_struct.visual
        if enable_cameras is not None:
            self._visual = self._visual or enable_cameras
        # sapien_env.py:321-327 + 899-907: the first reset seeds main/episode RNG with 2022+i BEFORE `_load_scene` runs, so
        # tasks can draw per-env geometry from `_batched_episode_rng` while building (peg_insertion_side.py:114-120)
        self._set_main_rng([2022 + i for i in range(num_envs)])
        self._set_episode_rng([2022 + i for i in range(num_envs)], np.arange(num_envs))
        self._control_mode_arg, self._fused_arg = control_mode, fused
        if render_mode not in self.SUPPORTED_RENDER_MODES + (None,):
            raise NotImplementedError(f"Unsupported render mode {render_mode}.")
        self.render_mode = render_mode
        # sapien_env.py:91-95,215-216: rebuild the scene every `reconfiguration_freq` resets (0 = never)
        self.reconfiguration_freq = int(reconfiguration_freq) if reconfiguration_freq is not None else 0
        self._reconfig_counter = 0
        self.scene = None
        self._reconfigure(dict())
        # sapien_env.py:321-327: main RNG seeds 2022+i, first reset
        self._set_main_rng([2022 + i for i in range(num_envs)])
        self._elapsed_steps[:] = 0
        self.reset(seed=[2022 + i for i in range(num_envs)])

    # ------------------------------------------------------------------ scene (re)build (sapien_env.py:725-770 `_reconfigure`)
    def _reconfigure(self, options: dict):
        """Build the sub-scene prototype from the task's description hooks, compile it and create the world; an existing world is
        destroyed first.  Tasks that draw geometry while building (per-env sizes from `_batched_episode_rng`) get new draws."""
        num_envs = self.num_envs
        if self.scene is not None and hasattr(self.scene.world, "close"):
            self.scene.world.close()
        self.scene_desc = SceneDesc(num_envs, self.sim_params)
        self._load_agent_desc()
        self._load_scene_desc()
        self.cm = self.scene_desc.compile()
        if self._world_factory is None:
            from ..backend import World
            world = World(self.cm, self._requested_device)
        else:
            world = self._world_factory(self.cm)
        self.scene = Scene(world, self.cm, self.scene_desc)
        self.device = self.scene.device
        self._elapsed_steps = torch.zeros(num_envs, dtype=torch.int32, device=self.device)
        self._hidden_objects = []
        self._after_build()
        self.agent.set_control_mode(self._control_mode_arg)
        self.single_action_space_low, self.single_action_space_high = self.agent.action_bounds()
        self.action_dim = self.single_action_space_low.shape[0]
        self._orig_single_action_space = SimpleNamespace(shape=(self.action_dim,), low=self.single_action_space_low, high=self.single_action_space_high)
        # `sapien.Device` of the simulation (sapien_env.py:1111 asks `.is_cuda()`): this backend is the GPU simulation by construction
        self._sim_device = Device(f"cuda:{self.device.index or 0}")
        self._sensors = self._setup_sensors() if self._visual else {}
        self._human_render_cameras = None    # created on the first render_rgb_array()
        self._last_obs = None
        self._fused = None
        # the fused control step serves the flat-state observation directly; tasks that can rebuild their observation dict from the
        # fused state vector (`_obs_from_fused`) use it under the visual modes too (B2S_FUSED_VISUAL=0 keeps the torch path there; tests/test_gpu_env.py compares the two)
        fused_visual = self._visual and hasattr(self, "_obs_from_fused") and os.environ.get("B2S_FUSED_VISUAL", "1") not in ("", "0")
        if self._fused_arg and self._world_factory is None and (self._obs_mode == "state" or fused_visual):
            self._fused = self._setup_fused_step()
        self._epilogue_runner = None
        if self._fused is None and self.graph_epilogue and self._world_factory is None and os.environ.get("B2S_GRAPH_EPILOGUE", "1") not in ("", "0"):
            self._epilogue_runner = _GraphedEpilogue(self)
        self._state_version += 1
        self._reconfig_counter = self.reconfiguration_freq

    # ------------------------------------------------------------------ task hooks (same names as the reference)
    def _load_agent_desc(self):
        raise NotImplementedError

    def _load_scene_desc(self):
        raise NotImplementedError

    def _after_build(self):
        raise NotImplementedError

    def _initialize_episode(self, env_idx: torch.Tensor, options: dict):
        pass

    def evaluate(self) -> dict:
        return {}

    def _get_obs_extra(self, info: dict) -> dict:
        return {}

    def compute_dense_reward(self, obs, action, info):
        raise NotImplementedError

    def compute_normalized_dense_reward(self, obs, action, info):
        raise NotImplementedError

    def compute_sparse_reward(self, obs, action, info):
        if "success" in info:
            if "fail" in info:
                return info["success"].float() - info["fail"].float()
            return info["success"].float()
        if "fail" in info:
            # sapien_env.py:693 writes `-info["fail"]`, which torch refuses for a bool tensor; the documented intent (-1 on failure)
            return -info["fail"].float()
        return torch.zeros(self.num_envs, device=self.device)

    # ------------------------------------------------------------------ properties
    @property
    def obs_mode(self):
        return self._obs_mode

    @property
    def reward_mode(self):
        return self._reward_mode

    @property
    def control_mode(self) -> str:
        """sapien_env.py `control_mode`: the agent's active control mode."""
        return self.agent.control_mode

    @property
    def elapsed_steps(self):
        return self._elapsed_steps

    @property
    def control_freq(self):
        return self._control_freq

    @property
    def sim_freq(self):
        return self._sim_freq

    @property
    def gpu_sim_enabled(self):
        return True

    # ------------------------------------------------------------------ RNG (sapien_env.py:980-1021)
    def _set_main_rng(self, seed):
        if seed is None:
            if self._main_seed is not None:
                return
            seed = np.random.RandomState().randint(2**31, size=(self.num_envs,))
        if not np.iterable(seed):
            seed = [seed]
        self._main_seed = list(seed)
        self._main_rng = np.random.RandomState(self._main_seed[0])
        if len(self._main_seed) == 1 and self.num_envs > 1:
            self._main_seed = self._main_seed + np.random.RandomState(self._main_seed[0]).randint(2**31, size=(self.num_envs - 1,)).tolist()
        self._batched_main_rng = U.BatchedRNG.from_seeds(self._main_seed)

    def _set_episode_rng(self, seed, env_idx):
        if seed is not None or self._enhanced_determinism:
            env_idx_np = env_idx.cpu().numpy() if isinstance(env_idx, torch.Tensor) else np.asarray(env_idx)
            if seed is None:
                self._episode_seed[env_idx_np] = self._batched_main_rng[env_idx_np].randint(2**31)
            else:
                if not np.iterable(seed):
                    seed = [seed]
                self._episode_seed = np.asarray(seed, dtype=np.int64)
                if len(self._episode_seed) == 1 and self.num_envs > 1:
                    self._episode_seed = np.concatenate((self._episode_seed, np.random.RandomState(self._episode_seed[0]).randint(2**31, size=(self.num_envs - 1,))))
            if seed is not None or self._batched_episode_rng is None:
                self._batched_episode_rng = U.BatchedRNG.from_seeds(self._episode_seed)
            else:
                self._batched_episode_rng[env_idx_np] = U.BatchedRNG.from_seeds(self._episode_seed[env_idx_np])
            self._episode_rng = self._batched_episode_rng[0]

    # ------------------------------------------------------------------ reset (sapien_env.py:857-978)
    def reset(self, seed: Union[None, int, Sequence[int]] = None, options: Optional[dict] = None):
        options = dict() if options is None else options
        if "env_idx" in options:
            env_idx = U.to_tensor(options["env_idx"], self.device, dtype=torch.int64).long()
        else:
            env_idx = torch.arange(0, self.num_envs, device=self.device)
        reconfigure = bool(options.get("reconfigure", False)) or (self._reconfig_counter == 0 and self.reconfiguration_freq != 0)
        if reconfigure and len(env_idx) != self.num_envs:
            raise RuntimeError("Cannot do a partial reset and reconfigure the environment. You must do one or the other.")   # sapien_env.py:903
        self._set_main_rng(seed)
        if reconfigure:  # sapien_env.py:909-916
            self._set_episode_rng(seed if seed is not None else self._batched_main_rng.randint(2**31), env_idx)
            with torch.random.fork_rng(devices=[self.device] if self.device.type == "cuda" else []):
                torch.manual_seed(int(self._episode_seed[0]))
                self._reconfigure(options)
            env_idx = env_idx.to(self.device)
            self._set_episode_rng(self._episode_seed, env_idx)   # again, so that what follows does not depend on the draws of the build
        else:
            self._set_episode_rng(seed, env_idx)
        self._state_version += 1  # invalidates per-state caches of derived poses (tasks may memoise them between fetches)
        if len(env_idx) == self.num_envs and hasattr(self.scene.world, "check_overflow"):
            self.scene.world.check_overflow()   # full resets are off the per-step path: report dropped contacts / rows loudly here
        self.scene._reset_mask = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.scene._reset_mask[env_idx] = True
        self.scene._reset_all = False
        self._elapsed_steps[env_idx] = 0
        self._clear_sim_state()
        if self.reconfiguration_freq != 0:
            self._reconfig_counter -= 1
        self.agent.reset()
        if "reset_to_env_states" in options:
            self.set_state_dict(options["reset_to_env_states"]["env_states"], env_idx)
        else:
            if seed is not None or self._enhanced_determinism:
                with torch.random.fork_rng(devices=[self.device] if self.device.type == "cuda" else []):
                    torch.manual_seed(int(self._episode_seed[0]))
                    self._initialize_episode(env_idx, options)
            else:
                self._initialize_episode(env_idx, options)
        self.scene._reset_mask = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        self.scene._reset_all = True
        # sapien_env.py:956-960: apply everything, refresh link poses, fetch
        self.scene._gpu_apply_all()
        self.scene._gpu_fetch_all()
        self.agent.controller_reset(env_idx)
        info = self.get_info()
        obs = self.get_obs(info)
        info["reconfigure"] = reconfigure
        self._last_obs = obs
        return obs, info

    def _clear_sim_state(self):
        """sapien_env.py:1023-1036: zero the velocities of the sub-scenes being reset."""
        for actor in self.scene.actors.values():
            if actor.px_body_type == "dynamic":
                actor.set_linear_velocity(torch.zeros(3, device=self.device))
                actor.set_angular_velocity(torch.zeros(3, device=self.device))
        for art in self.scene.articulations.values():
            art.set_qvel(torch.zeros(art.dof, device=self.device))
        self.scene._gpu_apply_all()

    # ------------------------------------------------------------------ step (sapien_env.py:1042-1132)
    def step(self, action):
        if self._fused is not None:
            return self._step_fused(action)
        action = self._step_action(action)
        self._elapsed_steps += 1
        self._state_version += 1
        if self._epilogue_runner is not None:
            info, core, reward, terminated = self._epilogue_runner.run(action)
        else:
            info, core, reward, terminated = self._epilogue(action)
        obs = self._obs_from_core(core)
        self._last_obs = obs
        return obs, reward, terminated, torch.zeros(self.num_envs, dtype=torch.bool, device=self.device), info

    # the part of `step` after the simulation (sapien_env.py:1056-1071): evaluate -> observation (without the pictures) -> reward ->
    # termination.  Pure tensor code over the persistent C-ABI buffers, no host synchronisation: tasks that set `graph_epilogue` have it
    # captured once into a CUDA graph and replayed (one launch instead of a few hundred small ones).
    graph_epilogue = False

    def _epilogue(self, action):
        info = self.get_info()
        if self._obs_mode == "none":
            core = dict()
        else:
            core = dict(agent=self._get_obs_agent(), extra=self._get_obs_extra(info))
            if self._obs_mode == "state" or self.obs_mode_struct.state:
                core = dict(state=U.flatten_state_dict(core))
        reward = self.get_reward(obs=core, action=action, info=info)
        if "success" in info:
            terminated = torch.logical_or(info["success"], info["fail"]) if "fail" in info else info["success"].clone()
        else:
            terminated = info["fail"].clone() if "fail" in info else torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        return info, core, reward, terminated

    def _obs_from_core(self, core):
        """`get_obs` from the pictureless part computed by `_epilogue` (the cameras are rendered here, outside the captured part)."""
        if self._obs_mode == "none":
            return dict()
        if self._obs_mode == "state":
            return core["state"]
        if self._obs_mode == "state_dict":
            return core
        if "state" in core:  # the flat state vector goes after the sensor entries (sapien_env.py:535-544 pops agent / extra, then adds it)
            obs = self._add_sensor_obs(dict())
            obs["state"] = core["state"]
            return obs
        return self._add_sensor_obs(dict(core))

    def _setup_fused_step(self):
        """task hook: return a fused-step handle (backend.create_pick_task) or None to use the torch path."""
        return None

    def _step_fused(self, action):
        """One C-ABI call = controller + substeps + evaluate/reward/obs kernels (include/b200sim.h b2s_pick_task_step);
        produces exactly what the torch path below produces (tests/test_gpu_env.py::test_fused_step_matches_python_path)."""
        if action is not None:
            if isinstance(action, np.ndarray):
                action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
            elif not isinstance(action, torch.Tensor):
                raise TypeError(type(action))
            action = action.to(device=self.device, dtype=torch.float32)
            if action.shape == (self.action_dim,):
                action = action[None].expand(self.num_envs, self.action_dim)   # the torch path broadcasts one action over the sub-scenes
            if tuple(action.shape) != (self.num_envs, self.action_dim):
                # the controller kernel reads actions[env * n_action + col] for every sub-scene: a wrong shape would read out of bounds
                raise ValueError(f"action must have shape {(self.num_envs, self.action_dim)} (or {(self.action_dim,)}), got {tuple(action.shape)}")
            action = action.contiguous()
        self.scene._gpu_apply_all()   # pending setter writes (set_qpos, set_pose, ...) reach the simulation state before the step
        f = self._fused
        self.scene.world.pick_task_step(f["handle"], action, self._sim_steps_per_control, f["obs"], f["reward"], f["flags"], self._elapsed_steps)
        self._state_version += 1
        # Like the torch path and the reference, every step hands out FRESH tensors: one clone of the packed outputs; the views below
        # share that clone, not the persistent kernel output buffers (which the next step overwrites).
        fl = f["flags"].clone()
        vec = f["obs"].clone()
        info = dict(elapsed_steps=self._elapsed_steps.clone(), success=fl[:, 0], is_obj_placed=fl[:, 1], is_robot_static=fl[:, 2], is_grasped=fl[:, 3])
        obs = vec if self._obs_mode == "state" else self._visual_obs_from_fused(vec, info)
        self._last_obs = obs
        return obs, f["reward"].clone(), fl[:, 4], torch.zeros(self.num_envs, dtype=torch.bool, device=self.device), info

    def supports_device_autoreset(self) -> bool:
        """True when `step_autoreset` is available: the fused control step is active and the task registered its episode
        initialisation with the backend."""
        return self._fused is not None and bool(self._fused.get("autoreset"))

    def enable_time_limit(self, max_episode_steps: int) -> bool:
        """The TimeLimit the vector wrapper enforces (registration.py:160-168), handed to the device-side auto-reset."""
        if not self.supports_device_autoreset():
            return False
        self._fused["max_episode_steps"] = int(max_episode_steps)
        return True

    def step_autoreset(self, action, ignore_terminations: bool = False):
        """`ManiSkillVectorEnv.step` for the fused task family with the auto-reset done on the device (include/b200sim.h
        b2s_pick_task_autoreset): no `dones.any()` host sync, no python-side partial reset.  Returns
        (obs, reward, terminated, truncated, info) where info always carries `final_info`, `final_observation` and the masks
        `_final_info` / `_final_observation` / `_elapsed_steps` (= done; all-false on steps where nothing finished -- the reference adds
        the keys only on such steps, which needs the host to look at `dones`).  Un-seeded resets draw from the torch CUDA generator."""
        if action is not None:
            if isinstance(action, np.ndarray):
                action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
            elif not isinstance(action, torch.Tensor):
                raise TypeError(type(action))
            action = action.to(device=self.device, dtype=torch.float32)
            if action.shape == (self.action_dim,):
                action = action[None].expand(self.num_envs, self.action_dim)
            if tuple(action.shape) != (self.num_envs, self.action_dim):
                raise ValueError(f"action must have shape {(self.num_envs, self.action_dim)} (or {(self.action_dim,)}), got {tuple(action.shape)}")
            action = action.contiguous()
        f = self._fused
        w = self.scene.world
        self.scene._gpu_apply_all()
        w.pick_task_step(f["handle"], action, self._sim_steps_per_control, f["obs"], f["reward"], f["flags"], self._elapsed_steps)
        self._state_version += 1
        visual = self._obs_mode != "state"
        if visual:
            self._sensors.capture()                      # the finished state, for final_observation
        rew = f["reward"].clone()
        elapsed = self._elapsed_steps.clone()
        rand = torch.rand((self.num_envs, 24), dtype=torch.float32, device=self.device)
        w.pick_task_autoreset(f["handle"], f["obs"], f["reward"], f["flags"], self._elapsed_steps, rand, f["final_obs"], f["done"], ignore_terminations,
                              f.get("max_episode_steps", 0))
        fl = f["flags"].clone()      # the finished step's flags, with the time-limit truncation the reset kernel decided
        done = f["done"].clone()
        if visual:
            self._sensors.keep_final(f["done"])          # pictures of the finished sub-scenes -> final buffers (masked copy)
            self._sensors.capture(env_mask=f["done"])    # re-render only what was reset
        vec = f["obs"].clone()
        final_vec = f["final_obs"].clone()
        info = dict(elapsed_steps=self._elapsed_steps.clone(), success=fl[:, 0], is_obj_placed=fl[:, 1], is_robot_static=fl[:, 2], is_grasped=fl[:, 3])
        final_info = dict(elapsed_steps=elapsed, success=fl[:, 0], is_obj_placed=fl[:, 1], is_robot_static=fl[:, 2], is_grasped=fl[:, 3])
        if visual:
            obs = self._visual_obs_from_fused(vec, info, render=False)
            final_obs = self._visual_obs_from_fused(final_vec, final_info, render=False, final=True)
        else:
            obs, final_obs = vec, final_vec
        info.update(final_info=final_info, final_observation=final_obs, _final_info=done, _final_observation=done, _elapsed_steps=done)
        self._last_obs = obs
        terminated = torch.zeros_like(fl[:, 4]) if ignore_terminations else fl[:, 4]
        return obs, rew, terminated, fl[:, 5], info

    def _step_action(self, action):
        if action is not None:
            if isinstance(action, np.ndarray):
                action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
            elif isinstance(action, torch.Tensor):
                action = action.to(self.device)
            else:
                raise TypeError(type(action))
            if action.shape == (self.action_dim,):
                action = action[None]
            self.agent.set_action(action)
            self.scene._gpu_apply_all()  # px.gpu_apply_articulation_target_position (sapien_env.py:1118-1121)
        self._before_control_step()
        cls = type(self)
        if cls._before_simulation_step is BaseEnv._before_simulation_step and cls._after_simulation_step is BaseEnv._after_simulation_step:
            # `for _ in range(self._sim_steps_per_control): scene.step()` + `_gpu_fetch_all()` as one fused launch
            self.scene.step(self._sim_steps_per_control, BUF_ALL)
        else:  # a task hooks into the individual simulation steps (sapien_env.py:1123-1128): one launch per step, then the fetch
            for _ in range(self._sim_steps_per_control):
                self._before_simulation_step()
                self.scene.step(1, 0)
                self._after_simulation_step()
            self.scene._gpu_fetch_all()
        self._after_control_step()
        return action

    def _before_simulation_step(self):
        """task hook, called before every simulation step of a control step (sapien_env.py:1473-1476)."""

    def _after_simulation_step(self):
        """task hook, called after every simulation step of a control step."""

    def _before_control_step(self):
        pass

    def _after_control_step(self):
        pass

    # ------------------------------------------------------------------ obs / info / reward (sapien_env.py:501-700)
    def get_info(self):
        info = dict(elapsed_steps=self._elapsed_steps.clone())
        info.update(self.evaluate())
        return info

    def _get_obs_agent(self):
        return self.agent.get_proprioception()

    def _get_obs_state_dict(self, info):
        return dict(agent=self._get_obs_agent(), extra=self._get_obs_extra(info))

    def get_obs(self, info=None, unflattened: bool = False):
        """sapien_env.py:501-533.  `unflattened` returns the raw nested dict (what the reward functions are handed in `step`)."""
        if info is None:
            info = self.get_info()
        if self._obs_mode == "none":
            return dict()
        if self._obs_mode in ("state", "state_dict"):
            obs = self._get_obs_state_dict(info)
        else:  # visual modes (sapien_env.py:535-625): agent + extra + sensor data / params
            obs = self._add_sensor_obs(dict(agent=self._get_obs_agent(), extra=self._get_obs_extra(info)))
        return obs if unflattened else self._flatten_raw_obs(obs)

    def _flatten_raw_obs(self, obs):
        """sapien_env.py:535-544: "state" flattens everything; a visual mode with the `state` flag replaces `agent` and `extra` by one
        flat `state` vector next to the sensor entries."""
        if self._obs_mode == "state":
            return U.flatten_state_dict(obs)
        if self.obs_mode_struct.state and isinstance(obs, dict) and "agent" in obs:
            obs["state"] = U.flatten_state_dict(dict(agent=obs.pop("agent"), extra=obs.pop("extra")))
        return obs

    def _add_sensor_obs(self, obs, render: bool = True, final: bool = False):
        """sapien_env.py:525-532: camera parameters + the requested textures, as a point cloud under obs mode "pointcloud"."""
        obs["sensor_param"] = self.get_sensor_params()
        obs["sensor_data"] = self._get_obs_sensor_data(render=render, final=final)
        return sensor_data_to_pointcloud(obs) if self.obs_mode_struct.pointcloud else obs

    def _visual_obs_from_fused(self, vec, info, render: bool = True, final: bool = False):
        """Visual-mode observation (same structure as `get_obs`) with the agent / extra entries -- or, under the `state` flag, the flat
        state vector that replaces them -- taken from the fused state vector.  render=False: the pictures are already taken;
        final=True: the pictures kept for the finished sub-scenes (`CameraSensors.keep_final`)."""
        if self.obs_mode_struct.state:
            obs = self._add_sensor_obs(dict(), render=render, final=final)
            obs["state"] = vec
            return obs
        return self._add_sensor_obs(self._obs_from_fused(vec, info), render=render, final=final)

    def _sensor_configs(self):
        """task hook: list of dict(uid, pose(7), width, height, fov, near, far, mount(link/actor name or None))."""
        return []

    def _setup_sensors(self):
        """sapien_env.py:771-843 `_setup_sensors` + scene.py:1087-1106: one camera group for all sensor cameras."""
        from ..render import CameraSensors, camera_desc
        cams = []
        for c in self._sensor_configs():
            c = dict(c)
            c.update(self._sensor_overrides.get(c["uid"], {}))
            row = -1
            if c.get("mount") is not None:
                art, link = c["mount"]
                row = self.cm.link_rows[art][link]
            cams.append(camera_desc(c["uid"], c["pose"], c["width"], c["height"], c["fov"], c["near"], c["far"], row))
        if not cams:
            raise NotImplementedError("this task defines no sensor cameras")
        # the sensors of the observation write only the textures the mode delivers (rgb 3 B + depth 2 B + segmentation 2 B per pixel) unless a
        # texture needs the raw targets (position) or the caller asks for them (`sensor_outputs="raw"`: get_picture_cuda stays available)
        m = self.obs_mode_struct
        outputs = 3
        if self._sensor_outputs == "compact" or (self._sensor_outputs == "auto" and m.visual and not m.position and not m.pointcloud):
            outputs = (4 if m.rgb else 0) | (8 if m.depth else 0) | (16 if m.segmentation else 0)
        return CameraSensors(self.scene.world, self.cm, cams, outputs=outputs or 3)

    def _get_obs_sensor_data(self, render: bool = True, final: bool = False):
        """sapien_env.py:578-625: hidden objects are simply absent from the sensor render-shape table (the reference
        teleports them away and back, actor.py:176-201), then update_render + take_picture on the camera group."""
        if render:
            self._sensors.capture()
        m = self.obs_mode_struct
        return self._sensors.get_obs(rgb=m.rgb, depth=m.depth, segmentation=m.segmentation, position=m.position, final=final)

    # ------------------------------------------------------------------ render modes (sapien_env.py:1369-1439)
    def _human_render_camera_configs(self):
        """task hook: cameras for `render_rgb_array` (same dict layout as `_sensor_configs`)."""
        return []

    def _make_camera_group(self, configs, overrides=None, include_hidden=False):
        from ..render import CameraSensors, camera_desc
        cams = []
        for c in configs:
            c = dict(c)
            c.update((overrides or {}).get(c["uid"], {}))
            row = -1
            if c.get("mount") is not None:
                owner, name = c["mount"]          # (articulation, link) or ("actor", actor name)
                row = self.cm.actor_rows[name] if owner == "actor" else self.cm.link_rows[owner][name]
            cams.append(camera_desc(c["uid"], c["pose"], c["width"], c["height"], c["fov"], c["near"], c["far"], row))
        return CameraSensors(self.scene.world, self.cm, cams, include_hidden=include_hidden)

    def render_rgb_array(self, camera_name: Optional[str] = None):
        """[num_envs, H, W, 3] uint8 from the human render cameras (all of them tiled, or the one named); hidden objects are shown."""
        if self._human_render_cameras is None:
            cfgs = self._human_render_camera_configs()
            if not cfgs:
                return None
            self._human_render_cameras = self._make_camera_group(cfgs, include_hidden=True)
        self._human_render_cameras.capture()
        data = self._human_render_cameras.get_obs(rgb=True, depth=False, segmentation=False)
        images = [v["rgb"] for k, v in data.items() if camera_name is None or k == camera_name]
        if not images:
            return None
        return images[0] if len(images) == 1 else tile_images(images)

    def get_sensor_images(self):
        """sapien_env.py:567-569: what the sensors currently see, as displayable images per sensor and texture."""
        if not self._sensors:
            self._sensors = self._setup_sensors()
        self._sensors.capture()
        data = self._sensors.get_obs(rgb=True, depth=True, segmentation=False)
        return {uid: camera_observations_to_images(d) for uid, d in data.items()}

    def render_sensors(self):
        return tile_images([img for d in self.get_sensor_images().values() for img in d.values()])

    def render_all(self):
        images = []
        human = self.render_rgb_array()
        if human is not None:
            images.append(human)
        images += [img for d in self.get_sensor_images().values() for img in d.values()]
        return tile_images(images)

    def render(self):
        if self.render_mode is None:
            raise RuntimeError("render_mode is not set.")
        if self.render_mode == "rgb_array":
            return self.render_rgb_array()
        if self.render_mode == "sensors":
            return self.render_sensors()
        if self.render_mode == "all":
            return self.render_all()
        raise NotImplementedError(f"Unsupported render mode {self.render_mode}.")

    def get_sensor_params(self):
        return self._sensors.get_params(self.scene.world.body_view()) if self._sensors else {}

    def get_reward(self, obs, action, info):
        if self._reward_mode == "sparse":
            return self.compute_sparse_reward(obs, action, info)
        if self._reward_mode == "dense":
            return self.compute_dense_reward(obs, action, info)
        if self._reward_mode == "normalized_dense":
            return self.compute_normalized_dense_reward(obs, action, info)
        return torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ state (sapien_env.py:1272-1325)
    def get_state_dict(self):
        return self.scene.get_sim_state()

    def set_state_dict(self, state, env_idx=None):
        self.scene.set_sim_state(state, env_idx)
        self.scene._gpu_apply_all()
        self.scene._gpu_fetch_all()
        self._state_version += 1

    def get_state(self):
        return U.flatten_state_dict(self.get_state_dict())

    def close(self):
        w = getattr(self.scene, "world", None)
        if w is not None and hasattr(w, "close"):
            w.close()
