"""Where a timepoint's wall time goes outside the engine's graph: ccb_begin_timepoint / ccb_ingest_device / the offline
phase (ccb_offline + export + Python objects), each bracketed by a device synchronize."""
import logging
import sys
import time

import torch

sys.path.insert(0, ".")
from chronoclust_b200.hddstream import HDDStream
from chronoclust_b200.synth import CONFIGS, config_params, gen

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
N, D, T, Cn, seed, eps, pi = CONFIGS[name]
N = int(N * scale)
Xs = gen(N, D, T, Cn, seed)
h = HDDStream(config_params(name), logging.getLogger("q"))
h.dataset_dimensionality = D
h._ensure_handle(D)
Xd = [torch.from_numpy(x).cuda() for x in Xs]
a = torch.empty(N, dtype=torch.int32, device="cuda")
s = torch.empty(N, dtype=torch.uint8, device="cuda")
for rep in range(3):
    h.reset()
    line = []
    for t in range(T):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.ingest_device(Xd[t].data_ptr(), N, D, t, a.data_ptr(), s.data_ptr(), run_offline=False)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        h.offline_clustering(t)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        line.append(f"t{t} online {1e3*(t1-t0):.2f} offline {1e3*(t2-t1):.2f}")
    print(f"rep {rep}: " + " | ".join(line))
