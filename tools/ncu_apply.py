"""ncu target: a few Gram-apply launches at the C3 shape with a forced tile shape.
usage: SLM_FORCE_APPLY_SHAPE=<id> python tools/ncu_apply.py <K>"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import get_engine  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 104
eng = get_engine()
p, F = 4096, 5
pa = eng.padded_cols(p)
G = torch.randn(F, pa, pa, dtype=torch.float64, device=eng.device)
ldz = (K + 7) // 8 * 8
Z = torch.randn(F, p, ldz, dtype=torch.float64, device=eng.device)
GZ = torch.empty_like(Z)
Karr = (ctypes.c_int32 * F)(*([K] * F))
for _ in range(4):
    eng.lib.slm_gram_apply(eng.h, eng._ptr(G), pa * pa, pa, p, F, Karr, eng._ptr(Z), ldz, eng._ptr(GZ), eng.stream)
torch.cuda.synchronize()
