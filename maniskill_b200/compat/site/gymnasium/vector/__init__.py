"""gymnasium.vector: the VectorEnv base class (0.29 signature) and utils.batch_space."""
from . import utils  # noqa: F401
from .utils import batch_space


class VectorEnv:
    metadata: dict = {}
    spec = None
    render_mode = None
    closed = False

    def __init__(self, num_envs: int, observation_space, action_space):
        self.num_envs = num_envs
        self.is_vector_env = True
        self.observation_space = batch_space(observation_space, n=num_envs)
        self.action_space = batch_space(action_space, n=num_envs)
        self.single_observation_space = observation_space
        self.single_action_space = action_space

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, actions):
        raise NotImplementedError

    def close_extras(self, **kwargs):
        pass

    def close(self, **kwargs):
        if self.closed:
            return
        self.close_extras(**kwargs)
        self.closed = True

    @property
    def unwrapped(self):
        return self

    def __del__(self):
        if not getattr(self, "closed", True):
            self.close()


class _VectorEnvOfCopies(VectorEnv):
    def __init__(self, env_fns, **kwargs):
        raise NotImplementedError("process-based vector environments are not part of this gymnasium subset")


SyncVectorEnv = AsyncVectorEnv = _VectorEnvOfCopies
