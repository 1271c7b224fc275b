"""One named column of a data frame (reference: src/anml/data/component.py:41-95).

Behaviour kept from the reference, including the quirk that a missing column
with a ``default_value`` is *added to the caller's frame* (component.py:84-87,
pinned by the reference's tests/data/test_component.py:45-49).
"""
from __future__ import annotations

from collections.abc import Iterable
from typing import Any, List, Optional

from pandas import DataFrame

from .frame import as_frame
from .validator import Validator


class Component:
    """Validate, fetch and cache ``df[key]``.

    Parameters
    ----------
    key : column name.
    validators : checks run on the column at ``attach``; ``None`` means none.
    default_value : used to create the column when the frame lacks it;
        ``None`` makes a missing column a ``KeyError``.
    """

    def __init__(self, key: str, validators: Optional[List[Validator]] = None, default_value: Optional[Any] = None):
        self.key = key
        self.validators = validators
        self.default_value = default_value
        self._value = None

    @property
    def key(self) -> str:
        return self._key

    @key.setter
    def key(self, key: str):
        if not isinstance(key, str):
            raise TypeError("Key of the component must be a string.")
        self._key = key

    @property
    def validators(self) -> List[Validator]:
        return self._validators

    @validators.setter
    def validators(self, validators: Optional[List[Validator]]):
        if validators is None:
            self._validators = []
            return
        if not isinstance(validators, Iterable):
            raise TypeError("Validators must be a list of validator.")
        validators = list(validators)
        if any(not isinstance(v, Validator) for v in validators):
            raise TypeError("Validators must be a list of validator.")
        self._validators = validators

    @property
    def value(self):
        """Cached column (``None`` before ``attach`` / after ``clear``); read-only."""
        return self._value

    def attach(self, df: DataFrame, skip=()) -> None:
        """Fetch, validate and cache the column.  ``skip`` names validator classes the
        caller will run itself (``Parameter.attach`` runs ``NoNans`` / ``Positive`` on the
        device after the upload instead of scanning the column on the host)."""
        df = as_frame(df)
        if self.key not in df:
            if self.default_value is None:
                raise KeyError(f"Dataframe doesn't contain column {self.key}.")
            df[self.key] = self.default_value
        column = df[self.key].to_numpy()
        for check in self.validators:
            if type(check) in skip:
                continue
            check(self.key, column)
        self._value = column

    def clear(self) -> None:
        self._value = None

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(key={self.key})"
