"""CPU oracle for the UnScene3D hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``unscene3d_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
use it, and there only as the checker / the timed CPU baseline.

PARITY UNPINNED at the MinkowskiEngine boundary: the reference ships no sparse-conv source and no
golden vectors; MinkowskiEngine (≈0.5.4, un-pinned `git clone`,
/root/reference/.devcontainer/Dockerfile:50-51) is absent from the container.  The oracle restates
ME's published semantics and is pinned instead by
  * dense-grid known-answer tests against torch.nn.functional.conv3d / conv_transpose3d /
    avg_pool3d (tests/test_oracle_dense_kat.py),
  * golden vectors produced by running the UNMODIFIED reference model files
    (/root/reference/models/*.py) on top of this oracle (tests/golden/, script
    tests/golden/make_golden.py),
  * the reference's own pure-torch functions where they run here unmodified (matcher, NCut).
"""
