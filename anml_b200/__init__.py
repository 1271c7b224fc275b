"""anml_b200 -- B200-native likelihood evaluation behind anml's Python API.

Sub-packages mirror the reference layout (``anml.data``, ``anml.prior``,
``anml.getter``, ``anml.variable``, ``anml.parameter``); ``anml_b200.model``
adds the thin objective/gradient/Hessian layer that consumes them.  All
design-matrix work runs in ``lib/libanml_b200.so`` (hand-written sm_100a CUDA).
"""
__version__ = "0.1.0"
