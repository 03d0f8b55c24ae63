"""CPU oracle for the satpy hot path (TEST INFRASTRUCTURE ONLY).

This package is a plain numpy / scipy / C restatement of the reference's
algorithms for: ABI/AHI/SEVIRI/VIIRS calibration, ``get_cos_sza`` +
``sunzen_corr_cos``, nearest-neighbour search + gather, bilinear and EWA
resampling.  Each function cites the reference file:line it follows.

It exists to CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
Nothing under ``satpy_b200/`` imports it; the product path has no CPU fallback
and fails loudly when the CUDA library is missing.

Parity status of each part (what pins it) is listed in ``DESIGN.md`` §3:
calibration, sun-zenith and the 2x2 nearest case are pinned by the reference's
own golden vectors (``tests/test_oracle_golden.py``); bilinear and EWA are
"parity unpinned" (the reference tests mock them out / have none) and are
restated from the published pyresample algorithm.
"""
