"""sapien.sensor (import only: the stereo-depth sensor simulation is outside the hot path)."""


class StereoDepthSensorConfig:
    def __init__(self):
        self.rgb_resolution = (1920, 1080)
        self.ir_resolution = (1280, 720)
        self.min_depth = 0.2
        self.max_depth = 10.0
        self.trans_pose_l = self.trans_pose_r = None


class StereoDepthSensor:
    def __init__(self, *a, **kw):
        raise NotImplementedError("the stereo depth sensor simulation is not available on the b200sim backend")
