"""Minimal `dacite`: from_dict(data_class, data, config=Config(strict=...)) for nested dataclasses with plain field types -- what
mani_skill/envs/sapien_env.py:265 does with SimConfig."""
from __future__ import annotations

import dataclasses
import typing
from dataclasses import dataclass


@dataclass
class Config:
    strict: bool = False
    check_types: bool = True
    cast: list = dataclasses.field(default_factory=list)


class DaciteError(Exception):
    pass


class UnexpectedDataError(DaciteError):
    def __init__(self, keys):
        super().__init__(f'can not match {sorted(keys)} to any data class field')
        self.keys = keys


def from_dict(data_class, data, config: Config = None):
    config = config or Config()
    hints = typing.get_type_hints(data_class)
    fields = {f.name: f for f in dataclasses.fields(data_class)}
    extra = set(data.keys()) - set(fields.keys())
    if config.strict and extra:
        raise UnexpectedDataError(extra)
    kwargs = {}
    for name, f in fields.items():
        if name not in data:
            continue
        value = data[name]
        tp = hints.get(name, f.type)
        if dataclasses.is_dataclass(tp) and isinstance(value, dict):
            value = from_dict(tp, value, config)
        kwargs[name] = value
    return data_class(**kwargs)
