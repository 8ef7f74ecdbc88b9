"""Import-level stand-in (mani_skill/vector/wrappers/sb3.py subclasses VecEnv)."""


class VecEnv:
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space


VecEnvObs = VecEnvStepReturn = VecEnvIndices = object
