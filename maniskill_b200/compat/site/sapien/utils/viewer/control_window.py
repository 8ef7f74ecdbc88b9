"""sapien.utils.viewer.control_window (import only: sapien_env.py:13)."""


class ControlWindow:
    pass
