from . import Scene  # noqa: F401
