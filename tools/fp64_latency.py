"""Single-warp latency probes on the GPU (cycles per dependent step): DADD, DMUL, DFMA, the replay pattern of
k_bs_chain_p (LDS, chained DADD, STS in place), FADD."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from chronoclust_b200 import _lib

L = _lib.lib()
sink = torch.zeros(8, dtype=torch.float64, device="cuda:0")
fl = C.c_double(0)
for _ in range(3):
    _lib.check(L.ccb_fp64_peak(0, None, 2, 0, 1, sink.data_ptr(), C.byref(fl)))
    torch.cuda.synchronize()
v = sink.cpu().tolist()
print("cycles per dependent step: DADD %.1f  DMUL %.1f  DFMA %.1f  replay(LDS+DADD+STS, 8-unrolled) %.1f  FADD %.1f" % tuple(v[1:6]))
