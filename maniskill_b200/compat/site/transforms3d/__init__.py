"""The part of `transforms3d` the reference uses (euler, quaternions; quaternions are wxyz)."""
from . import euler, quaternions  # noqa: F401
