"""TEST INFRASTRUCTURE ONLY — CPU restatement of the MedicalSeg VNet hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and there only
as the checker or as the timed CPU baseline, never as the thing shipped.

Parity status
-------------
* VNet / losses / optimizer (``vnet_oracle.py``): **parity unpinned**.  The reference ships no
  tests, golden vectors or fixtures for this path and its arithmetic lives in
  PaddlePaddle (``paddlepaddle-gpu>=2.2.0``, unpinned, not installable offline).  The restatement
  follows the reference's call sites line by line and encodes Paddle's documented defaults.
* Preprocess (``preprocess_oracle.py``): **pinned** against the reference's own
  ``tools/preprocess_utils/{values,geometry}.py`` executed in the build container
  (``tests/golden/make_golden.py`` -> ``tests/golden/preprocess_*.npz``).
"""
