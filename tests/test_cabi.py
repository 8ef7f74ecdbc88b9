"""CPU: the C-ABI library builds, loads and exports every symbol include/b200sim.h declares; creating a world without a
GPU fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    so = os.path.join(ROOT, "maniskill_b200", "libb200sim.so")
    if not os.path.exists(so):
        g.build()
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "b200sim.h")).read()
    declared = set(re.findall(r"\b(b2s_[a-z_]+)\s*\(", hdr)) - {"b2s_env_step_fused"}
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), name


def test_no_gpu_no_world():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from maniskill_b200.backend import World
    from maniskill_b200.scenes import pick_cube_scene
    with pytest.raises(RuntimeError):
        World(pick_cube_scene(1).compile())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "maniskill_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#include\s+[\"<][^\">]*oracle", r"libb2s_oracle", r"b2o_"):
                    assert not re.search(pat, src, re.M), f"{f} uses the oracle ({pat})"
