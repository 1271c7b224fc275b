"""Frames a ``Component`` can attach to.

The reference reads pandas frames (``df[key].to_numpy()``, data/component.py:84-91).
A ``pyarrow.Table`` is accepted as well (Arrow -> device ingest):
a null-free float64 column that is one chunk comes out as a NumPy *view* of the Arrow
buffer and goes straight into the pinned-lane upload, nulls become NaN (and are caught
by ``NoNans``), and the default-valued columns a ``Component`` adds to the caller's frame
live beside the immutable table.
"""
from __future__ import annotations

import numpy as np


class _Column:
    __slots__ = ("_values",)

    def __init__(self, values: np.ndarray):
        self._values = values

    def to_numpy(self) -> np.ndarray:
        return self._values


class ArrowFrame:
    """The slice of the pandas interface ``Component.attach`` uses, over a ``pyarrow.Table``."""

    def __init__(self, table):
        self.table = table
        self._names = set(table.column_names)
        self._added = {}

    def __len__(self) -> int:
        return self.table.num_rows

    def __contains__(self, key) -> bool:
        return key in self._added or key in self._names

    def __getitem__(self, key) -> _Column:
        if key in self._added:
            return _Column(self._added[key])
        if key not in self._names:
            raise KeyError(key)
        col = self.table.column(key)
        if col.num_chunks == 1:
            values = col.chunk(0).to_numpy(zero_copy_only=False)  # a view unless there are nulls / a conversion
        else:
            values = col.to_numpy()  # chunks are concatenated (one copy)
        return _Column(values)

    def __setitem__(self, key, value) -> None:
        arr = np.asarray(value)
        self._added[key] = np.full(len(self), value) if arr.ndim == 0 else arr


def as_frame(df):
    """pandas frames pass through; a ``pyarrow.Table`` is wrapped (once)."""
    if isinstance(df, ArrowFrame):
        return df
    if hasattr(df, "column_names") and hasattr(df, "num_rows") and hasattr(df, "column"):
        return ArrowFrame(df)
    return df
