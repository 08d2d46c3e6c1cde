#!/usr/bin/env python3
"""Wall time of the whole drop-in entry point chronoclust_b200.app.run(...) -- CSV parsing, scaling, the hot path,
tracking, result / per-cell output files and the pickled program images included ("report the app.run wall
time beside the hot-path numbers").  The inputs are the C2 generator's timepoints written as CSV files.

    python tools/app_wall.py [scale]        # scale = fraction of C2's 1e6 cells per timepoint (default 0.2)
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chronoclust_b200 import app  # noqa: E402
from chronoclust_b200.synth import CONFIGS, config_params, gen  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
N, D, T, Cn, seed, eps, pi = CONFIGS["C2"]
N = int(N * scale)
cfg = config_params("C2")
with tempfile.TemporaryDirectory() as tmp:
    files = []
    t0 = time.perf_counter()
    for t, X in enumerate(gen(N, D, T, Cn, seed)):
        f = os.path.join(tmp, f"d{t}.csv")
        np.savetxt(f, X, delimiter=",", header=",".join(f"m{d}" for d in range(D)), comments="", fmt="%.17g")
        files.append(f)
    t_write = time.perf_counter() - t0
    res = {}
    for normalise in (False, True, False):  # (the first pass also pays the one-time costs: CUDA context, library load)
        out = os.path.join(tmp, f"out{len(res)}_{int(normalise)}")
        os.makedirs(out)
        t0 = time.perf_counter()
        app.run(data=files, output_directory=out, normalise_data=normalise, param_beta=cfg["beta"], param_delta=cfg["delta"],
                param_epsilon=cfg["epsilon"], param_lambda=cfg["lambda"], param_k=cfg["k"], param_mu=cfg["mu"],
                param_pi=cfg["pi"], param_omicron=cfg["omicron"], param_upsilon=cfg["upsilon"])
        wall = time.perf_counter() - t0
        nres = sum(1 for _ in open(os.path.join(out, "result.csv"))) - 1
        name = ("first_run_" if not res else "") + ("normalise" if normalise else "no_normalise")
        res[name] = {"seconds": wall, "cells_per_s": N * T / wall, "result_rows": nres,
                     "output_bytes": sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)
                                         if f.endswith(".csv"))}
    print(json.dumps({"metric": "app.run wall time, everything included", "cells": N * T, "runs": res,
                      "input_csv_write_s": t_write, "host_cores": os.cpu_count(),
                      "workload": f"C2 at scale {scale}: {N} cells x {D} markers x {T} timepoints as CSV files"}))
