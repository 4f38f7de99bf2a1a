"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker / the timed CPU baseline.
The product (``tpnet_b200``) never imports this package and raises if its CUDA
library is missing.
"""
