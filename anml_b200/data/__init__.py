"""Data ingest: pull validated columns out of a pandas frame (host side)."""
from .component import Component
from .example import DataExample
from .prototype import DataPrototype
from .validator import NoNans, Positive, Unique, Validator

__all__ = ["Component", "DataExample", "DataPrototype", "NoNans", "Positive", "Unique", "Validator"]
