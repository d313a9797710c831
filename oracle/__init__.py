"""CPU oracle for the Sparse2Dense hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``sparse2dense_b200/`` imports this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` leg and
``--impl reference``).  See ``oracle/s2d_oracle.c`` for the per-function citations
into the reference and for the pinned / unpinned status of each piece.
"""
