"""Host-side mirrors of the three tasks BASELINE.json names (PickCube-v1, PegInsertionSide-v1, OpenCabinetDrawer-v1): what the GPU box --
which has no checkout of the reference -- runs for the benchmark and the `-m gpu` tests.  Every other task of the reference is not
re-typed here: its own module runs unmodified on this backend through the `sapien` shim (maniskill_b200/compat), see
tests/test_reference_unmodified.py."""
from ..registration import register_env
from .base_env import BaseEnv
from .pick_cube import PickCubeEnv

register_env("PickCube-v1", max_episode_steps=50)(PickCubeEnv)
from .peg_insertion_side import PegInsertionSideEnv

register_env("PegInsertionSide-v1", max_episode_steps=100)(PegInsertionSideEnv)
from .open_cabinet_drawer import OpenCabinetDrawerEnv

register_env("OpenCabinetDrawer-v1", max_episode_steps=100)(OpenCabinetDrawerEnv)
