"""CPU oracle for the anml likelihood-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``anml_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline
legs of ``bench.py`` do, and there only as the checker / the timed baseline.
"""
