"""CPU oracle for the pybader hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``pybader_b200/`` imports this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.
"""
