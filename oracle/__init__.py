"""CPU oracle for the deformation hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there
only as the checker (or as the reported CPU baseline), never as the thing shipped.

Provenance
----------
* ``jacobian_oracle``  -- PINNED: restates ``eval_reg_with_jacobian.py:62-78`` of the
  reference, and is checked against golden vectors produced by executing those very
  lines of the reference file (``tests/golden/make_jacobian_golden.py``).
* ``interp_oracle`` / ``torch_oracle`` -- PARITY UNPINNED: the arithmetic lives in the
  un-vendored third-party packages voxelmorph@52dd120f, neurite@c7bb05d5,
  pystrum@8cd5c483 on TensorFlow 2.7 (reference ``README.md:35-42``), none of which can be
  installed here (no network, no wheels).  The restatement follows their published
  algorithm as recalled in SURVEY.md Appendix A; the reference has no tests or golden
  vectors for this path.  It is anchored by oracle-independent analytic known-answer
  tests and an independent float64 cross-check (scipy ``map_coordinates``).
"""
