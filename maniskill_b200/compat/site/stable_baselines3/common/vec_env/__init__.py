from . import base_vec_env  # noqa: F401
