"""CPU oracle for the KP2DTiny perception hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``nano_vs_slam_b200/`` may import this
package: it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` have an independent
CPU statement of what the reference computes.

Every function cites the reference ``file:line`` it restates (paths relative
to the upstream repository root).  The restatement is *functional*: it works on
a plain ``state_dict`` (``{name: tensor}``) with the reference's key names
instead of ``nn.Module`` trees.

Parity pinning:
  * model forward / post_processing: PINNED.  ``oracle/gen_golden.py`` imports
    the real reference package from ``/root/reference`` in the build container,
    runs it on seeded inputs + spread-init weights and stores the outputs in
    ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle
    against those vectors (and against the live reference when it is present).
  * matcher: the arithmetic lives in OpenCV (``cv2.BFMatcher``), which is not
    part of the reference tree; the restated ``goodMatchesOneToOne`` is checked
    against ``cv2`` outputs stored in ``tests/golden/matcher_*.npz``.
  * retrieval: the arithmetic lives in faiss-cpu 1.9.0 (``IndexFlatL2``), which
    is neither vendored nor installable here -> **parity unpinned** for that
    row; the oracle restates the published IndexFlatL2 semantics
    (exact squared L2, k smallest ascending, int64 labels).
"""
