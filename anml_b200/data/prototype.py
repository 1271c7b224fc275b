"""A named bundle of components (reference: src/anml/data/prototype.py:26-57)."""
from __future__ import annotations

from typing import Dict

from pandas import DataFrame

from .component import Component


class DataPrototype:
    """Expose every entry of ``components`` as an attribute of the same name and
    attach / clear them together."""

    def __init__(self, components: Dict[str, Component]):
        self.components = components
        for name, comp in components.items():
            setattr(self, name, comp)

    @property
    def components(self) -> Dict[str, Component]:
        return self._components

    @components.setter
    def components(self, components: Dict[str, Component]):
        for name, comp in components.items():
            if not isinstance(name, str):
                raise TypeError("Components key must be a string.")
            if not isinstance(comp, Component):
                raise TypeError(f"Components {name} value must be a instance of Component")
        self._components = components

    def attach(self, df: DataFrame) -> None:
        for name in self.components:
            getattr(self, name).attach(df)

    def clear(self) -> None:
        for name in self.components:
            getattr(self, name).clear()

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(components={self.components})"
