"""Timeline of the block-speculative engine, round by round, on the CUDA-graph path (debug build only: csrc/debug.h,
ccb_debug_trace).  Every record carries the control-block fields of the round and the START time of every kernel of that
round (globaltimer), so the critical path of a round -- gaps included -- can be read off without a profiler.

    python tools/trace_rounds.py C2 1.0 [chunk] [--eps 0.04] [--tps 2] [--out gpurun_out/trace.npz]
"""
import ctypes as C
import logging
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from chronoclust_b200 import _lib as _l0, build as _b0

_l0.SO_PATH = _b0.build(debug=True)
from chronoclust_b200.hddstream import HDDStream
from chronoclust_b200.synth import CONFIGS, config_params, gen

KERNELS = ["begin", "spec", "need", "nearest", "merge", "spec_o", "tilecnt", "pscan", "pscatter", "chain_p", "derive_p",
           "verify_p", "olist", "chain_o", "derive_o", "verify_o", "decide", "commit_rows", "commit_cells", "finish",
           "chain_p_end", "longest_chain", "-", "-"]
FIELDS = ["kind", "pos", "Bcur", "Beff_in", "it", "nneed_in", "npend", "nh", "no", "m0", "upgrade", "pclean_next", "m_commit",
          "Mp", "Mo0", "hnew0", "nneed", "Beff", "t_end"]


def opt(name, default, cast=float):
    if name in sys.argv:
        i = sys.argv.index(name)
        v = cast(sys.argv[i + 1])
        del sys.argv[i:i + 2]
        return v
    return default


eps_over = opt("--eps", None)
tps = opt("--tps", 2, int)
out = opt("--out", "gpurun_out/trace_rounds.npz", str)
detail = opt("--detail", 12, int)
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N, D, T, Cn, seed, eps, pi = CONFIGS[name]
N = int(N * scale)
Xs = gen(N, D, T, Cn, seed)
prm = config_params(name)
if eps_over is not None:
    prm["epsilon"] = eps_over
h = HDDStream(prm, logging.getLogger("q"), chunk=chunk)
h.dataset_dimensionality = D
h._ensure_handle(D)
fn = _l0.lib().ccb_debug_trace
fn.restype, fn.argtypes = C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64]
Xd = [torch.from_numpy(x).cuda() for x in Xs[:tps]]
a = torch.empty(N, dtype=torch.int32, device="cuda")
s = torch.empty(N, dtype=torch.uint8, device="cuda")
buf = np.zeros((1 << 14, 64), dtype=np.int64)
saved = {}
for rep in range(2):  # the second pass is the warm one
    h.reset()
    for t in range(tps):
        fn(h._h, buf.ctypes.data_as(C.c_void_p), 0)  # rewind
        h.ingest_device(Xd[t].data_ptr(), N, D, t, a.data_ptr(), s.data_ptr())
        torch.cuda.synchronize()
        n = fn(h._h, buf.ctypes.data_as(C.c_void_p), buf.shape[0])
        saved[f"t{t}"] = buf[:min(n, buf.shape[0])].copy()
np.savez_compressed(out, **saved)
cn = np.zeros(16, dtype=np.int64)
fc = _l0.lib().ccb_debug_counters
fc.restype, fc.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]
fc(h._h, cn.ctypes.data_as(C.c_void_p), 1)
print("debug counters:", cn.tolist())

for t in range(tps):
    r = saved[f"t{t}"]
    rounds = r[r[:, 0] < 2]
    blocks = r[r[:, 0] == 2]
    t_first = r[0, 20 + 0] if len(r) else 0
    print(f"== t{t}: {len(blocks)} blocks, {len(rounds)} rounds, {(r[-1, 18] - t_first) / 1e6:.2f} ms first begin -> last finish")
    # per-round durations: from the previous record's end (or the block's begin) to this decide's end
    prev_end = None
    rows = []
    for rec in r:
        kind = rec[0]
        ts = rec[20:60]
        if kind == 2:
            start = prev_end if prev_end is not None else ts[17]
            rows.append(("commit", rec, (rec[18] - start) / 1e3))
        else:
            start = prev_end if (prev_end is not None and rec[4] > 0) else ts[0]
            rows.append(("round", rec, (rec[18] - start) / 1e3))
        prev_end = rec[18]
    tot_round = sum(d for k, _, d in rows if k == "round")
    tot_commit = sum(d for k, _, d in rows if k == "commit")
    light = [d for k, rec, d in rows if k == "round" and rec[6] > 0 and rec[20 + 9] == 0]
    full = [d for k, rec, d in rows if k == "round" and rec[20 + 9] != 0]
    print(f"   rounds {tot_round / 1e3:.2f} ms (full {len(full)}: {np.sum(full) / 1e3:.2f} ms, mean {np.mean(full) if full else 0:.0f} us; "
          f"light {len(light)}: {np.sum(light) / 1e3:.2f} ms, mean {np.mean(light) if light else 0:.0f} us), commits {tot_commit / 1e3:.2f} ms")
    # kernel-start offsets inside full rounds, relative to the round's start, averaged over first rounds (it == 0) and later ones
    for label, sel in (("first rounds (it = 0)", lambda rec: rec[4] == 0 and rec[20 + 9] != 0),
                       ("later full rounds", lambda rec: rec[4] > 0 and rec[20 + 9] != 0),
                       ("light rounds", lambda rec: rec[4] > 0 and rec[20 + 9] == 0)):
        acc = {}
        cnt = 0
        pe = None
        for rec in r:
            if rec[0] < 2 and sel(rec):
                ts = rec[20:60]
                start = ts[0] if rec[4] == 0 else pe
                if start:
                    cnt += 1
                    for k, nm in enumerate(KERNELS[:20]):
                        if ts[k] > 0 and (rec[4] == 0 or k >= 3):
                            acc.setdefault(nm, []).append((ts[k] - start) / 1e3)
                    if ts[20] > 0:
                        acc.setdefault("chain_p_end", []).append((ts[20] - start) / 1e3)
                    acc.setdefault("decide_end", []).append((rec[18] - start) / 1e3)
            pe = rec[18]
        if cnt:
            print(f"   {label}: n={cnt}; mean start offset (us): " +
                  "  ".join(f"{nm} {np.mean(v):.0f}" for nm, v in acc.items()))
    print("   first blocks:")
    shown = 0
    for k, rec, d in rows:
        if shown >= detail * 6:
            break
        shown += 1
        if k == "commit":
            print(f"      COMMIT pos={rec[1]} m={rec[9]} created={rec[3]} upgrade={rec[10]} Mp={rec[13]} Mo0={rec[14]}  {d:.0f} us")
        else:
            print(f"      round it={rec[4]} B={rec[2]} Beff={rec[3]}->{rec[17]} nneed={rec[5]}->{rec[16]} npend={rec[6]} nh={rec[7]} "
                  f"(new from {rec[15]}) no={rec[8]} m0={rec[9]} longest={rec[20 + 21]} {'COMMIT' if rec[0] == 1 else ('light next' if rec[11] else 'full next')}  {d:.0f} us")
