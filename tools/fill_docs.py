#!/usr/bin/env python3
"""Generates DESIGN.md / README.md from tools/templates/*.in: the @PLACEHOLDER@ fields are filled from the measured files
under profiles/ (final passes of round 2), so that every number quoted in the documents is the one in the committed
evidence.  Edit the templates, then re-run.

    python tools/fill_docs.py <1-GPU tag, e.g. r2z> <multi-GPU tag, e.g. r2y> [<tag of the 1-GPU line of the scaling tables>]
"""
import json
import re
import sys

one, multi = sys.argv[1], sys.argv[2]
P = "profiles/"
b = json.load(open(f"{P}{one}_bench_c2.json"))
app = json.load(open(f"{P}{one}_app_wall_c2.json"))["runs"]
tp = open(f"{P}{one}_tp_wall_c2_graph.log").read().strip().splitlines()[-1]
t = [float(x) for x in re.findall(r"t\d (\d+\.\d+)ms", tp)]
one_multi = sys.argv[3] if len(sys.argv) > 3 else multi
bn = {N: json.load(open(f"{P}{multi if N > 1 else one_multi}_bench_c2_{N}gpu.json")) for N in (1, 2, 4, 8)}
sw = json.load(open(f"{P}{multi}_sweep64_8gpu.json"))
if not bn[1].get("offline_c4"):  # (a quick 1-GPU line without the C4 leg: the final pass has it)
    bn[1]["offline_c4"] = b["offline_c4"]

M = lambda v: f"{v / 1e6:.1f}"
sf, rd = b["serial_floor"], b["roofline_distance"]
c4rows, c4line = [], []
t1 = bn[1]["offline_c4"]["ms"]
for N in (1, 2, 4, 8):
    c = bn[N]["offline_c4"]
    s = c["stage_ms_rank0"]
    ex = s.get("csr_rowinfo", 0.0) + s.get("csr_fill_allreduce", 0.0) + s.get("allgather_submask", 0.0)
    c4rows.append(f"| {N} | {c['ms']:.1f} | {t1 / c['ms']:.2f} | {t1 / c['ms'] / N:.2f} | {s['neighbours']:.1f} / {s['subspace']:.1f} / "
                  f"{s['weighted']:.1f} / {ex:.2f} / {s['clusters']:.2f} |")
    c4line.append(f"{c['ms']:.1f} ms on {N}")
rep = []
for N in (1, 2, 4, 8):
    j = bn[N]
    rep.append(f"  | {N} | {M(j['value'])} M / {M(j['e2e']['value'])} M cells/s | {j['value'] / bn[1]['value'] / N:.3f} / "
               f"{j['e2e']['value'] / bn[1]['e2e']['value'] / N:.3f} | {min(j['per_rank_ms_per_step']):.1f}–{max(j['per_rank_ms_per_step']):.1f} |")
reptable = "  | GPUs | value / e2e | efficiency | ms per step, fastest–slowest rank |\n  |---|---|---|---|\n" + "\n".join(rep)
sm = sw["seconds_per_run_min_median_max"]
sweep = (f"{sw['configs']} runs of 5e6 cells in **{sw['seconds']:.1f} s = {M(sw['value'])} M cells/s** on {sw['n_gpus']} GPUs "
         f"(runs per rank {sw.get('runs_per_rank')}; per run {sm[0]:.2f} s fastest, {sm[1]:.2f} s median, {sm[2]:.2f} s slowest).")
vals = {
    "VALUE": M(b["value"]), "MS": f"{b['ms_per_step']:.1f}", "E2E": M(b["e2e"]["value"]), "E2EP": M(b["e2e_pageable"]["value"]),
    "CPU": f"{b['cpu_baseline']['value'] / 1e6:.2f}",
    "APP": f"{app['no_normalise']['cells_per_s'] / 1e6:.2f}", "APP_S": f"{app['no_normalise']['seconds']:.1f}",
    "APPN": f"{app['normalise']['cells_per_s'] / 1e6:.2f}", "APPN_S": f"{app['normalise']['seconds']:.1f}",
    "T0": f"{t[0]:.1f}", "T1": f"{t[1]:.1f}", "T2": f"{t[2]:.1f}", "T3": f"{t[3]:.1f}", "T4": f"{t[4]:.1f}",
    "FLOOR": f"{sf['floor_ms_per_step']:.2f}", "CHAIN": f"{sf['measured_ms_per_step']:.1f}", "FRAC": f"{sf['frac']:.2f}",
    "K1": f"{rd['achieved']:.2f}", "K1F": f"{100 * rd['frac']:.1f}", "K1N": f"{100 * rd['frac_of_nofma']:.0f}",
    "PFMA": f"{rd['peak']:.1f}", "PNOFMA": f"{rd['peak_nofma']:.1f}",
    "C3": M(b["c3"]["value"]), "C3E": M(b["c3"]["e2e_pageable"]),
    "C4TABLE": "\n".join(c4rows), "C4LINE": ", ".join(c4line) + f" GPUs (efficiency {t1 / bn[8]['offline_c4']['ms'] / 8:.2f} at 8)",
    "REPTABLE": reptable,
    "REPLINE": " / ".join(M(bn[N]["value"]) for N in (1, 2, 4, 8)) + " M cells/s device-resident, "
               + " / ".join(M(bn[N]["e2e"]["value"]) for N in (1, 2, 4, 8)) + " M end to end",
    "SWEEP": sweep, "SWEEPLINE": f"{sw['seconds']:.1f} s for 64 runs of 5e6 cells = {M(sw['value'])} M cells/s",
}
for path in ("DESIGN.md", "README.md"):
    s = open(f"tools/templates/{path}.in").read()  # the documents are GENERATED: edit the templates, not the outputs
    for k, v in vals.items():
        s = s.replace(f"@{k}@", v)
    left = re.findall(r"@[A-Z0-9_]+@", s)
    open(path, "w").write(s)
    print(path, "unfilled:", left)
