"""ORACLE — test infrastructure, not product code.

CPU restatement of the reference's algorithm for the PCG hot path (``uibk/deep_preconditioning/cg.py``,
``test.py:61-68,100-109``, ``utils.py:15-43``, ``model.py:53-57``). Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and only as the checker or the timed CPU baseline — never from ``deeppreconditioning_b200``.

Pinning status
--------------
* ``oracle.pcg`` (the loop, ``cg.py:15-90``): pinned against the reference function itself, imported from
  ``/root/reference`` in the build container (``tests/test_oracle_vs_reference.py``; iteration counts equal on
  every case) and against golden vectors that run produced (``tests/golden/pcg_golden.json``, made by
  ``tests/golden/make_golden.py``).
* ``oracle.sparse.sparse_matvec_mul`` (``utils.py:15-43``): pinned by the reference's own known-answer test
  (``tests/test_utils.py:11-41``).
* ``oracle.sparse`` assembly (``test.py:61-68,100-105``): pinned against ``torch.Tensor.to_sparse_csr`` /
  dense arithmetic, which is literally what the reference executes.
* SpTRSV / level sets: the reference has none — pinned against scipy ``spsolve_triangular``.
* IC(0): ``ilupp`` 1.0.2 is absent — **parity unpinned** for values; pattern and the defining property
  ``(L L^T)_ij = A_ij`` on the pattern are tested.
* CNN values (spconv 2.3.8 absent): **parity unpinned**; structural properties of
  ``tests/test_model.py:31-42`` are tested.
"""
