"""CPU oracle for the DynamicScaler panorama tiling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dynamicscaler_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker /
the timed CPU baseline.

Parity status: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``
and committed as ``tests/golden/*.npz``.
"""
