"""TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy fp64 / complex128) of the TemGymCore hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``temgymcore_b200`` never imports it.
"""
