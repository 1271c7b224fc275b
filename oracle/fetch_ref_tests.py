"""Recipe: make the REFERENCE's own unit tests available where /root/reference does not exist.

The GPU box only receives this repository.  `tests/test_reference_suite_replay.py` replays the
reference's test-suite, unchanged, against this package (``anml`` aliased to ``anml_b200``); the files
that need a design matrix can only run there.  This script copies ``/root/reference/tests`` into
``oracle/_ref/tests`` -- git-ignored (no reference source enters the history) but not gpurun-ignored, so
it travels with the snapshot like a built ``.so``.  ``__graft_entry__.build()`` runs it.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/tests"
DST = os.path.join(HERE, "_ref", "tests")


def fetch() -> str | None:
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return DST


if __name__ == "__main__":
    print(fetch())
