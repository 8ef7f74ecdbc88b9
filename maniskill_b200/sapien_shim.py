"""A `sapien`-shaped module tree over this backend: the part of the `sapien` surface that exists here, under
the names the reference imports.

    from maniskill_b200 import sapien_shim as sapien
    sapien.Pose(p=[0, 0, 1]);  sapien.physx.PhysxGpuSystem(world);  sapien.physx.PhysxMaterial(1, 1, 0)
    sapien.ActorBuilder();  sapien.wrapper.articulation_builder.ArticulationBuilder();  sapien.render.RenderMaterial(base_color=[...])

`install()` registers the tree in `sys.modules` as `sapien`, `sapien.physx`, `sapien.render`, `sapien.wrapper.*` when no real `sapien` is
importable, so that `import sapien` in code written for the reference resolves here.  What the tree does not have (entities / components,
the URDF loader, the viewer, lights, textures) raises AttributeError on access -- nothing is silently stubbed.
"""
from __future__ import annotations

import sys
import types

from . import building as _b
from . import physx_shim as _px

_GPU_ENABLED = False
_ME = sys.modules[__name__]


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


class Device:
    """`sapien.Device("cuda" | "cuda:n" | "cpu")` (mani_skill/envs/utils/system/backend.py:60-91)."""

    def __init__(self, name: str):
        self.name = name
        kind, _, idx = name.partition(":")
        if kind not in ("cpu", "cuda"):
            raise ValueError(f"unknown device {name!r}")
        self.type, self.cuda_id = kind, (int(idx) if idx else 0)

    def is_cuda(self) -> bool:
        return self.type == "cuda"

    def is_cpu(self) -> bool:
        return self.type == "cpu"

    def __repr__(self):
        return f"Device({self.name!r})"


def _enable_gpu():
    global _GPU_ENABLED
    _GPU_ENABLED = True


Pose = _b.Pose
ActorBuilder = _b.ActorBuilder

physx = _module(
    __name__ + ".physx", PhysxGpuSystem=_px.PhysxGpuSystem, PhysxMaterial=_b.PhysxMaterial, enable_gpu=_enable_gpu,
    is_gpu_enabled=lambda: _GPU_ENABLED)
render = _module(__name__ + ".render", RenderMaterial=_b.RenderMaterial)
wrapper = _module(__name__ + ".wrapper")
wrapper.articulation_builder = _module(__name__ + ".wrapper.articulation_builder", ArticulationBuilder=_b.ArticulationBuilder, LinkBuilder=_b.LinkBuilder,
                                       JointRecord=_b.JointRecord)
wrapper.actor_builder = _module(__name__ + ".wrapper.actor_builder", ActorBuilder=_b.ActorBuilder, CollisionShapeRecord=_b.CollisionRecord,
                                VisualShapeRecord=_b.VisualRecord)


def install(force: bool = False) -> bool:
    """Register this tree as `sapien` in sys.modules.  Returns False (and does nothing) when a real `sapien` can be imported, unless
    `force`."""
    if not force:
        import importlib.util
        if "sapien" in sys.modules and sys.modules["sapien"] is not _ME:
            return False
        if "sapien" not in sys.modules and importlib.util.find_spec("sapien") is not None:
            return False
    sys.modules["sapien"] = _ME
    sys.modules["sapien.physx"] = physx
    sys.modules["sapien.render"] = render
    sys.modules["sapien.wrapper"] = wrapper
    sys.modules["sapien.wrapper.articulation_builder"] = wrapper.articulation_builder
    sys.modules["sapien.wrapper.actor_builder"] = wrapper.actor_builder
    return True
