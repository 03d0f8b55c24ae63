#!/usr/bin/env python
"""Path statistics of the fixed-grid search on config 5 (how many targets need the disc step, the lattice model,
the exact fall-back ...): DESIGN.md section 5 quotes them.  Needs a SCRATCH build of the library with the counters
compiled in -- they are not part of the shipped library:

    mkdir -p scratch/dbg && cd satpy_b200/csrc && for f in *.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 \
        -std=c++17 -Xcompiler -fPIC -DB200_DEBUG_COUNTERS -c $f -o ../../scratch/dbg/${f%.cu}.o; done && \
        nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/dbg/libsatpy_b200.so ../../scratch/dbg/*.o
    python scripts/path_counters.py          # on a B200
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from satpy_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "scratch", "dbg", "libsatpy_b200.so")
import torch  # noqa: E402

from satpy_b200 import demo  # noqa: E402
from satpy_b200.pipeline import ABIResamplePipeline  # noqa: E402

src = demo.make_area("goes_east_abi_f_1km"); dst = demo.make_area("global_eqc_500m")
counts = demo.abi_counts(src.shape, 7, "C02"); cal = demo.abi_calibration("C02")
lib = _lib.load()
pipe = ABIResamplePipeline(src, dst, 2000.0)
cd = torch.from_numpy(counts).cuda() if not isinstance(counts, torch.Tensor) else counts.cuda()
buf = (C.c_ulonglong * 16)()
lib.b200_debug_counters.argtypes = [C.c_void_p, C.c_int]
lib.b200_debug_counters(buf, 1)
out = pipe.run(cd, cal, demo.START_TIME)
lib.b200_debug_counters(buf, 1)
names = ["searched", "w>=1 (disc needed)", "model ok", "model: shortcut failed -> loop", "no-model loop", "exact fallback",
         "cand evals model loop", "cand evals no-model loop", "w>=1 with all 4 invalid", "no-model though 4 valid",
         "no valid pixel near", "leader outside radius"]
n = buf[0]
for i, nm in enumerate(names):
    print(f"{nm:36s} {buf[i]:14d}  {buf[i]/max(n,1):.4f}")
print("total target", out.numel(), "searched frac", n / out.numel(), "finite frac", float(torch.isfinite(out).float().mean()))
