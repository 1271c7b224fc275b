"""One design-matrix column and its priors (reference: src/anml/variable/main.py:57-184)."""
from __future__ import annotations

from typing import List, Optional, Tuple, Type, Union

import numpy as np
from numpy.typing import NDArray
from pandas import DataFrame

from ..data.component import Component
from ..data.validator import NoNans
from ..prior.main import Prior
from ..prior.utils import filter_priors, get_prior_type


class Variable:
    """``component``: a column name (wrapped into a ``Component`` with a
    ``NoNans`` validator) or a ready ``Component``; ``priors``: priors on this
    variable's coefficient(s)."""

    _prior_types: Tuple[Type, ...] = (Prior,)

    def __init__(self, component: Union[str, Component], priors: Optional[List[Prior]] = None):
        self.component = component
        self.priors = priors

    @property
    def component(self) -> Component:
        return self._component

    @component.setter
    def component(self, component: Union[str, Component]):
        if isinstance(component, str):
            component = Component(component, validators=[NoNans()])
        elif not isinstance(component, Component):
            raise TypeError("Variable input component has to be a string or an instance of Component.")
        self._component = component

    @property
    def priors(self) -> List[Prior]:
        return self._priors

    @priors.setter
    def priors(self, priors: Optional[List[Prior]]):
        priors = [] if priors is None else list(priors)
        if any(not isinstance(p, self._prior_types) for p in priors):
            raise TypeError("Variable input priors must be a list of instances of Prior.")
        self._priors = priors

    @property
    def size(self) -> Optional[int]:
        return 1

    def attach(self, df: DataFrame) -> None:
        self.component.attach(df)

    def get_design_mat(self, df: DataFrame) -> NDArray:
        """(n, 1) view of the column (main.py:99-114)."""
        self.attach(df)
        return self.component.value[:, np.newaxis]

    # hooks used by Parameter.attach: write this variable's columns straight into the
    # device-resident design matrix (no host hstack).  NoNans / Positive are not run on
    # the host here: Parameter.attach checks all uploaded columns in one device pass and
    # raises the same ValueError as the host validators.
    def _write_design(self, design, col0: int, df: DataFrame) -> None:
        from ..data.validator import NoNans, Positive

        self.component.attach(df, skip=(NoNans, Positive))
        design.set_columns(col0, self.component.value[:, np.newaxis])

    def _device_checks(self):
        """(wants NoNans, wants Positive) for the columns this variable uploaded."""
        from ..data.validator import NoNans, Positive

        kinds = {type(v) for v in self.component.validators}
        return NoNans in kinds, Positive in kinds

    def get_direct_prior_params(self, prior_type: str) -> NDArray:
        """(2, size) parameters of the single mat-less prior of this type, or the
        type's defaults repeated (main.py:116-151)."""
        direct = filter_priors(self.priors, prior_type, with_mat=False)
        cls = get_prior_type(prior_type)
        if not direct:
            return np.repeat(cls.default_params, self.size, axis=1)
        if len(direct) > 1:
            raise ValueError("Variable can only have one direct prior.")
        if direct[0].shape[1] != self.size:
            raise ValueError("Variable and prior size don't match.")
        return direct[0].params

    def get_linear_prior_params(self, prior_type: str) -> Tuple[NDArray, NDArray]:
        """Stacked ((2, m), (m, size)) of every prior of this type that has a
        linear map; empty arrays when there is none (main.py:153-184)."""
        linear = filter_priors(self.priors, prior_type, with_mat=True)
        if not linear:
            return np.empty((2, 0)), np.empty((0, self.size))
        if any(p.shape[1] != self.size for p in linear):
            raise ValueError("Variable and prior size don't match.")
        return np.hstack([p.params for p in linear]), np.vstack([p.mat for p in linear])

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(component={self.component}, priors={self.priors})"
