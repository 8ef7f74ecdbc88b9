"""Import-level stand-in for matplotlib (mani_skill/envs/tasks/tabletop/place_sphere.py imports pyplot at module level)."""
from . import animation, pyplot  # noqa: F401


def use(*a, **kw):
    pass
