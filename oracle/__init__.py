"""CPU oracle for the NDCN ODE-integrated graph-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ndcn_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker /
the CPU arm, never as the thing measured as "ours" or shipped.
"""
