"""TEST INFRASTRUCTURE — CPU oracle for the VoteNet hot path.  NOT product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package; the
product (``votenet_b200``) never does and has no CPU fallback.

* ``oracle.ops``   — numpy front-end (same Python signatures as the reference's ``tf_ops/*/tf_*.py`` wrappers) over
  ``oracle/_build/liboracle.so`` (the C restatement in ``oracle.c``), plus ``oracle.ops.ref`` — the same functions
  served by the REAL reference sources compiled unmodified into ``oracle/_ref/`` (when built).
* ``oracle.dense`` — torch-CPU fp32 restatement of the reference's TensorFlow/Tensorpack glue
  (``utils.py`` pointnet_sa_module / pointnet_fp_module, ``model.py`` voting / proposal / decode).
"""
