from . import control_window  # noqa: F401
