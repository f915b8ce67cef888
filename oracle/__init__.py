"""CPU oracle for the artifact-datastore hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``metaflow_b200`` (the product) may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference legs of
``bench.py`` do, and there only as the checker or as the timed CPU baseline.
"""
