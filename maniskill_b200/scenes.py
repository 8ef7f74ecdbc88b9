"""Scene prototypes for the stock tasks, restated from the reference's builders.

* table scene:   mani_skill/utils/scene_builder/table/scene_builder.py:20-66
* ground plane:  mani_skill/utils/building/ground.py:37-44
* Panda agent:   mani_skill/agents/robots/panda/panda.py:16-98 (urdf_config friction 2.0 on both fingers,
                 PD gains 1e3/1e2, force limit 100), gravity disabled on all robot links
                 (mani_skill/agents/base_agent.py:278-282)
* PickCube-v1:   mani_skill/envs/tasks/tabletop/pick_cube.py:79-104
"""
from __future__ import annotations

import numpy as np

from .model import (SHAPE_BOX, SHAPE_PLANE, SHAPE_SPHERE, ActorRec, ArticulationRec, SceneDesc, ShapeRec, SimParams,
                    load_robot, pose7)

TABLE_HEIGHT = 0.9196429
SQRT_HALF = float(np.sqrt(0.5))


def panda_articulation(name="panda", urdf="panda_v2", root_p=(-0.615, 0.0, 0.0)) -> ArticulationRec:
    robot = load_robot(urdf)
    drive = {}
    for i in range(1, 8):
        drive[f"panda_joint{i}"] = (1e3, 1e2, 100.0)
    drive["panda_finger_joint1"] = (1e3, 1e2, 100.0)
    drive["panda_finger_joint2"] = (1e3, 1e2, 100.0)
    art = ArticulationRec(name, robot, pose7(root_p), link_mu={"panda_leftfinger": 2.0, "panda_rightfinger": 2.0},
                          disable_gravity=True, drive=drive)
    art.link_patch = {"panda_leftfinger": 0.1, "panda_rightfinger": 0.1}
    return art


def add_table_scene(scene: SceneDesc):
    table = ActorRec(
        "table-workspace", "kinematic",
        [ShapeRec(SHAPE_BOX, pose7([0, 0, TABLE_HEIGHT / 2]), np.array([2.418 / 2, 1.209 / 2, TABLE_HEIGHT / 2]),
                  color=(0.62, 0.47, 0.33, 1.0))],
        pose7([-0.12, 0, -TABLE_HEIGHT], [SQRT_HALF, 0, 0, SQRT_HALF]))
    scene.add_actor(table)
    ground = ActorRec(
        "ground", "static",
        [ShapeRec(SHAPE_PLANE, pose7([0, 0, -TABLE_HEIGHT], [0.7071068, 0, -0.7071068, 0]), color=(0.45, 0.45, 0.45, 1.0))],
        pose7())
    scene.add_actor(ground)


def pick_cube_scene(n_envs: int, sim: SimParams | None = None, cube_half_size=0.02, goal_thresh=0.025) -> SceneDesc:
    scene = SceneDesc(n_envs, sim)
    scene.add_articulation(panda_articulation())
    add_table_scene(scene)
    scene.add_actor(ActorRec("cube", "dynamic", [ShapeRec(SHAPE_BOX, pose7(), np.array([cube_half_size] * 3), color=(1, 0, 0, 1))],
                             pose7([0, 0, cube_half_size])))
    scene.add_actor(ActorRec("goal_site", "kinematic",
                             [ShapeRec(SHAPE_SPHERE, pose7(), np.array([goal_thresh, 0, 0]), color=(0, 1, 0, 1), collide=False)],
                             pose7(), hidden=True))
    return scene


PANDA_REST_QPOS = np.array([0.0, np.pi / 8, 0, -np.pi * 5 / 8, 0, np.pi * 3 / 4, np.pi / 4, 0.04, 0.04])
