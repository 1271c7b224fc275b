"""Observation container (reference: src/anml/data/example.py:6-26).

``obs`` must be free of NaNs; ``obs_se`` defaults to 1.0 and must be positive.
It is the only "observation" container the reference has, so the model layer
(:mod:`anml_b200.model`) takes its observations through it.
"""
from .component import Component
from .prototype import DataPrototype
from .validator import NoNans, Positive


class DataExample(DataPrototype):
    def __init__(self, obs: str, obs_se: str):
        super().__init__(
            {
                "obs": Component(obs, [NoNans()]),
                "obs_se": Component(obs_se, [NoNans(), Positive()], default_value=1.0),
            }
        )
