"""Runs the UNMODIFIED reference python layer (haosulab/ManiSkill, `mani_skill`) on the b200sim backend -- SURVEY.md section 8(b) B1.

    import maniskill_b200.compat as compat
    compat.install()                       # `sapien` (+ the small pure-python packages the reference imports) resolve to compat/site/
    sys.path.insert(0, "<checkout of haosulab/ManiSkill>")
    import gymnasium as gym, mani_skill.envs
    env = gym.make("PickCube-v1", num_envs=4096, obs_mode="state")        # sim_backend="physx_cuda" by default for num_envs > 1

`install()` puts `compat/site` on `sys.path` -- behind the real site-packages, so anything genuinely installed (a real `gymnasium`,
`transforms3d`, ...) wins; only missing modules resolve to the stand-ins.  The `sapien` package in there is the b200sim-backed
implementation of the surface the reference uses (compat/site/sapien/__init__.py); the world it creates at `gpu_init()` comes from
`WORLD_FACTORY` (default: maniskill_b200.backend.World, i.e. CUDA through the C-ABI; tests substitute the host emulation).
"""
from __future__ import annotations

import importlib.util
import os
import sys

SITE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "site")

# callable(compiled_model, device_index) -> world with the interface of maniskill_b200.backend.World; None = the CUDA backend
WORLD_FACTORY = None


def install(force_sapien: bool = False) -> str:
    """Make `import sapien`, `import gymnasium`, ... resolvable.  Returns the directory that was added to sys.path."""
    if SITE not in sys.path:
        sys.path.append(SITE)   # last: real packages take precedence
    # a stub registered as `sapien` (a module object without a file: a test's mock) must not shadow the package; a really installed sapien is left alone
    mod = sys.modules.get("sapien")
    if mod is not None and not isinstance(getattr(mod, "__file__", None), str):
        for k in [k for k in sys.modules if k == "sapien" or k.startswith("sapien.")]:
            del sys.modules[k]
    return SITE


def _is_ours(name: str) -> bool:
    spec = importlib.util.find_spec(name)
    return spec is not None and spec.origin is not None and spec.origin.startswith(SITE)


def make_world(cm, device_index: int = 0):
    if WORLD_FACTORY is not None:
        return WORLD_FACTORY(cm, device_index)
    import torch
    from ..backend import World
    return World(cm, torch.device("cuda", device_index))
