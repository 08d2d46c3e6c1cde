"""Single-warp latency probes on the GPU (cycles per dependent step): DADD, DMUL, DFMA, the replay pattern of
k_bs_chain_p (LDS, chained DADD, STS in place), FADD."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from chronoclust_b200 import _lib

L = _lib.lib()
sink = torch.zeros(16, dtype=torch.float64, device="cuda:0")
fl = C.c_double(0)
for _ in range(3):
    _lib.check(L.ccb_fp64_peak(0, None, 2, 0, 1, sink.data_ptr(), C.byref(fl)))
    torch.cuda.synchronize()
v = sink.cpu().tolist()
print("cycles per dependent step: DADD %.1f  DMUL %.1f  DFMA %.1f  replay(LDS+DADD+STS, 8-unrolled) %.1f  FADD %.1f" % tuple(v[1:6]))
for _ in range(3):
    _lib.check(L.ccb_fp64_peak(0, None, 3, 0, 1, sink.data_ptr(), C.byref(fl)))
    torch.cuda.synchronize()
v = sink.cpu().tolist()
print("replay schedule of k_bs_chain_p, cycles per cell: DADD chain only %.1f | + addends from LDS %.1f | + STS of every version "
      "%.1f | same with two warps spinning on an mbarrier %.1f" % tuple(v[8:12]))
