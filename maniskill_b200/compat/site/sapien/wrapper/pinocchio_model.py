"""sapien.wrapper.pinocchio_model (CPU inverse kinematics of the CPU simulation path; import only on this backend)."""


class PinocchioModel:
    def __init__(self, *a, **kw):
        raise NotImplementedError("pinocchio models belong to the CPU simulation path; the GPU controllers use batched torch kinematics")
