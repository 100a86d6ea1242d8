"""CPU oracle — TEST INFRASTRUCTURE, not product code.

NumPy restatement of the reference's per-frame association path (tracker plugins under
/root/reference/plugins/track/*, the tracker wrappers under /root/reference/tracklab/wrappers/track/*
and the pre/post-processing they rely on). Every function cites the reference file:line it follows.

Rules (enforced by tests/test_layout.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import anything from ``oracle``;
  * ``tracklab_b200`` never imports it — the product path fails loudly without the CUDA library.

Pinning: the reference has no tests for this path, so the oracle is pinned against
OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by importing the unmodified
plugins from /root/reference (tests/golden/make_golden.py; fixtures committed under tests/golden/).
The third-party solvers the reference calls that are absent offline (lap 0.5.12) are replaced by
their published semantics on top of scipy (``oracle/ref_shims``, ``oracle/assign_np.py``) — that part
is "parity unpinned" against lap's own binaries and is anchored on uniqueness of the optimum.
"""
