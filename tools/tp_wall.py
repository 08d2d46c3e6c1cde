"""Per-timepoint wall time of the device-resident path without event timing (what bench.py's `value` sees)."""
import logging
import sys
import time

import torch

sys.path.insert(0, ".")
from chronoclust_b200.hddstream import HDDStream
from chronoclust_b200.synth import CONFIGS, config_params, gen

if "--debuglib" in sys.argv:  # experiment knobs (CCB_SLACK, ...) only exist in the debug build
    sys.argv.remove("--debuglib")
    from chronoclust_b200 import _lib as _l0, build as _b0
    _l0.SO_PATH = _b0.build(debug=True)
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
N, D, T, Cn, seed, eps, pi = CONFIGS[name]
N = int(N * scale)
Xs = gen(N, D, T, Cn, seed)
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h = HDDStream(config_params(name), logging.getLogger("q"), chunk=chunk)
h.dataset_dimensionality = D
h._ensure_handle(D)
Xd = [torch.from_numpy(x).cuda() for x in Xs]
a = torch.empty(N, dtype=torch.int32, device="cuda")
s = torch.empty(N, dtype=torch.uint8, device="cuda")
for rep in range(3):
    h.reset()
    prev = h.stats()
    line = []
    for t in range(T):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.ingest_device(Xd[t].data_ptr(), N, D, t, a.data_ptr(), s.data_ptr())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = h.stats()
        line.append(f"t{t} {dt*1e3:.1f}ms blocks={st['bsv_blocks']-prev['bsv_blocks']} rounds={st['bsv_rounds']-prev['bsv_rounds']} "
                    f"light={st['bsv_light_rounds']-prev['bsv_light_rounds']} "
                    f"cuts={st['bsv_cuts_unknown']-prev['bsv_cuts_unknown']}/{st['bsv_cuts_rounds']-prev['bsv_cuts_rounds']}/"
                    f"{st['bsv_cuts_capacity']-prev['bsv_cuts_capacity']} launches={st['kernel_launches']-prev['kernel_launches']}")
        prev = st
    print(f"rep {rep}: " + " | ".join(line) + f" | pdl={h.stats()['bsv_pdl']}")
