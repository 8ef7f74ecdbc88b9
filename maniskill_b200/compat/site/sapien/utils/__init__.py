"""sapien.utils: the interactive viewer is out of scope (needs a window system); the class exists so that imports resolve."""
from . import viewer  # noqa: F401


class Viewer:
    def __init__(self, *a, **kw):
        raise RuntimeError("the interactive sapien Viewer is not available on the b200sim backend (render_mode='human' needs a display)")
