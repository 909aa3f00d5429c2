"""Shared machinery for the ``<Kind>Builder.from_parameters(dict).build()`` pattern
(reference: crates/rscm-core/src/python/component.rs:19-47, `create_component_builder!`)."""

from __future__ import annotations

from .core import Component


class ComponentBuilder:
    KIND: int = 0
    TYPE_NAME: str = ""
    # (field name, default or None when required)
    FIELDS: tuple = ()

    def __init__(self, parameters: dict):
        self._parameters = dict(parameters)

    @classmethod
    def from_parameters(cls, parameters: dict):
        known = {n for n, _ in cls.FIELDS}
        unknown = set(parameters) - known
        if unknown:
            raise ValueError(f"{cls.TYPE_NAME}: unknown parameter(s) {sorted(unknown)}")
        missing = [n for n, d in cls.FIELDS if d is None and n not in parameters]
        if missing:
            # pythonize::depythonize of the parameter struct fails on a missing field
            raise ValueError(f"{cls.TYPE_NAME}: missing parameter(s) {missing}")
        return cls(parameters)

    def _value(self, name, default):
        return float(self._parameters.get(name, default))

    def build(self) -> Component:
        names = [n for n, _ in self.FIELDS]
        vals = [self._value(n, d) for n, d in self.FIELDS]
        return Component(self.KIND, self.TYPE_NAME, names, vals)
