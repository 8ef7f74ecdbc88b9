"""sapien.wrapper.articulation_builder: `LinkBuilder`, `ArticulationBuilder`, joint / mimic records (articulation_builder.py:49-112,
161-170 of the reference subclass these and read the record fields)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

import sapien
from sapien import physx

from .actor_builder import ActorBuilder


@dataclass
class JointRecord:
    joint_type: str = "undefined"   # fixed | revolute | revolute_unwrapped | continuous | prismatic | free | undefined
    limits: tuple = ()
    pose_in_parent: sapien.Pose = field(default_factory=sapien.Pose)
    pose_in_child: sapien.Pose = field(default_factory=sapien.Pose)
    friction: float = 0
    damping: float = 0
    name: str = ""


@dataclass
class MimicJointRecord:
    joint: str
    mimic: str
    multiplier: float
    offset: float


class LinkBuilder(ActorBuilder):
    def __init__(self, index, parent):
        super().__init__()
        self.parent = parent
        self.index = index
        self.joint_record = JointRecord()
        self.physx_body_type = "link"

    def set_joint_name(self, name):
        self.joint_record.name = name
        return self

    def set_joint_properties(self, type, limits, pose_in_parent=None, pose_in_child=None, friction=0, damping=0):
        self.joint_record.joint_type = type
        self.joint_record.limits = limits
        self.joint_record.pose_in_parent = pose_in_parent if pose_in_parent is not None else sapien.Pose()
        self.joint_record.pose_in_child = pose_in_child if pose_in_child is not None else sapien.Pose()
        self.joint_record.friction = friction
        self.joint_record.damping = damping
        return self

    def set_parent(self, parent):
        self.parent = parent
        return self

    def _check(self):
        valid = ("fixed", "revolute", "revolute_unwrapped", "continuous", "prismatic", "free", "undefined")
        assert self.joint_record.joint_type in valid, f"invalid joint type {self.joint_record.joint_type}"
        if self.joint_record.joint_type in ("revolute", "prismatic", "revolute_unwrapped"):
            assert np.asarray(self.joint_record.limits).size == 2, "a 1-dof joint needs one [lower, upper] limit pair"


class ArticulationBuilder:
    def __init__(self):
        self.link_builders: List[LinkBuilder] = []
        self.mimic_joint_records: List[MimicJointRecord] = []
        self.scene = None
        self.initial_pose = sapien.Pose()
        self.name = ""

    def set_scene(self, scene):
        self.scene = scene
        return self

    def set_initial_pose(self, pose):
        self.initial_pose = pose
        return self

    def set_name(self, name):
        self.name = name
        return self

    def create_link_builder(self, parent: Optional[LinkBuilder] = None):
        if self.link_builders:
            assert parent and parent in self.link_builders
        builder = LinkBuilder(len(self.link_builders), parent)
        self.link_builders.append(builder)
        return builder

    def build_entities(self, fix_root_link=None, name_prefix=""):
        entities, links = [], []
        for b in self.link_builders:
            b._check()
            b.physx_body_type = "link"
            entity = sapien.Entity()
            link = b.build_physx_component(links[b.parent.index] if b.parent else None)
            entity.add_component(link)
            if b.visual_records:
                entity.add_component(b.build_render_component())
            entity.name = b.name
            link.name = f"{name_prefix}{b.name}"
            j, r = link.joint, b.joint_record
            j.name, j.type = f"{name_prefix}{r.name}", r.joint_type
            j.pose_in_child, j.pose_in_parent = r.pose_in_child, r.pose_in_parent
            if j.type in ("revolute", "prismatic", "revolute_unwrapped"):
                j.limit = np.array(r.limits).flatten()
                j.set_drive_property(0, r.damping)
            if j.type == "continuous":
                j.limit = [-np.inf, np.inf]
                j.set_drive_property(0, r.damping)
            links.append(link)
            entities.append(entity)
        if fix_root_link is not None:
            entities[0].components[0].joint.type = "fixed" if fix_root_link else "undefined"
        entities[0].pose = self.initial_pose
        return entities

    def build(self, fix_root_link=None, name_prefix=""):
        assert self.scene is not None
        entities = self.build_entities(fix_root_link, name_prefix)
        articulation = entities[0].components[0].articulation
        articulation.pose = self.initial_pose
        for e in entities:
            self.scene.add_entity(e)
        return articulation
