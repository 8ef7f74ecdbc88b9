from ..registration import register_env
from .base_env import BaseEnv
from .pick_cube import PickCubeEnv

register_env("PickCube-v1", max_episode_steps=50)(PickCubeEnv)
from .peg_insertion_side import PegInsertionSideEnv

register_env("PegInsertionSide-v1", max_episode_steps=100)(PegInsertionSideEnv)
from .open_cabinet_drawer import OpenCabinetDrawerEnv

register_env("OpenCabinetDrawer-v1", max_episode_steps=100)(OpenCabinetDrawerEnv)
from .push_cube import PushCubeEnv

register_env("PushCube-v1", max_episode_steps=50)(PushCubeEnv)
from .stack_cube import StackCubeEnv

register_env("StackCube-v1", max_episode_steps=50)(StackCubeEnv)
from .pull_cube import PullCubeEnv

register_env("PullCube-v1", max_episode_steps=50)(PullCubeEnv)
from .lift_peg_upright import LiftPegUprightEnv

register_env("LiftPegUpright-v1", max_episode_steps=50)(LiftPegUprightEnv)
from .poke_cube import PokeCubeEnv

register_env("PokeCube-v1", max_episode_steps=50)(PokeCubeEnv)
from .roll_ball import RollBallEnv

register_env("RollBall-v1", max_episode_steps=80)(RollBallEnv)
from .place_sphere import PlaceSphereEnv

register_env("PlaceSphere-v1", max_episode_steps=50)(PlaceSphereEnv)
from .stack_pyramid import StackPyramidEnv

register_env("StackPyramid-v1", max_episode_steps=250)(StackPyramidEnv)
from .pull_cube_tool import PullCubeToolEnv

register_env("PullCubeTool-v1", max_episode_steps=100)(PullCubeToolEnv)
from .plug_charger import PlugChargerEnv

register_env("PlugCharger-v1", max_episode_steps=200)(PlugChargerEnv)
