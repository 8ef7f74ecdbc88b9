from . import actor_builder, articulation_builder, coacd, pinocchio_model, urdf_loader  # noqa: F401
