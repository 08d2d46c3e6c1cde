"""ncu capture window in STEADY STATE: timepoints 0..1 of the workload run unprofiled, then cudaProfilerStart() brackets
the first `cells` cells of timepoint 2 (stream launches, so every kernel of a block is an individual launch).

    ncu --profile-from-start off --set full -k regex:k_bs_chain_p -c 4 python tools/profile_window.py C2 0.2 40000
    python tools/profile_window.py k1      # the dense kernel-1 benchmark (1e6 cells x 4096 MCs x D=12), one launch
"""
import ctypes as C
import logging
import sys

import torch

sys.path.insert(0, ".")
from chronoclust_b200 import _lib
from chronoclust_b200.hddstream import HDDStream
from chronoclust_b200.synth import CONFIGS, config_params, gen

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
if name == "k1":
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    N, M = 1_000_000, 4096
    g = torch.Generator(device="cpu").manual_seed(5)
    X = torch.rand((N, D), dtype=torch.float64, generator=g).cuda()
    cen = torch.rand((M, D), dtype=torch.float64, generator=g).cuda()
    mask = torch.randint(0, 2 ** min(D, 62), (M,), dtype=torch.int64, generator=g).cuda()
    slot = torch.empty(N, dtype=torch.int32, device="cuda")
    dist = torch.empty(N, dtype=torch.float64, device="cuda")
    L = _lib.lib()
    for rep in range(3):
        if rep == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        _lib.check(L.ccb_nearest(0, None, X.data_ptr(), N, D, D, cen.data_ptr(), mask.data_ptr(), M, 4.0, slot.data_ptr(),
                                 dist.data_ptr()))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 40000
N, D, T, Cn, seed, eps, pi = CONFIGS[name]
N = int(N * scale)
Xs = gen(N, D, 3, Cn, seed)
h = HDDStream(config_params(name), logging.getLogger("q"), bsv_stream=1)
h.dataset_dimensionality = D
h._ensure_handle(D)
Xd = [torch.from_numpy(x).cuda() for x in Xs]
a = torch.empty(N, dtype=torch.int32, device="cuda")
s = torch.empty(N, dtype=torch.uint8, device="cuda")
for t in range(2):
    h.ingest_device(Xd[t].data_ptr(), N, D, t, a.data_ptr(), s.data_ptr())
torch.cuda.synchronize()
torch.cuda.profiler.start()
h.ingest_device(Xd[2].data_ptr(), min(cells, N), D, 2, a.data_ptr(), s.data_ptr(), run_offline=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(h.stats())
