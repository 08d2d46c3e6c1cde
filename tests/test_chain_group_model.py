"""CPU suite: the CONTESTED-group logic of the pcore replay (csrc/engine.cuh: fast / exact radius test, two cells per step,
eight cells per pass along predicted verdicts), modelled operation by operation in tests/proto/chain_group_model.c, on
microclusters grown to their radius limit (the regime where these paths run):

* the division-free fast test never contradicts the reference's radius test (utilities/mc_functions.py:45-56) when it
  decides -- and it does decide nearly always, otherwise it would be no shortcut;
* both schedules leave the verdicts and the state of the one-by-one replay, bit for bit, whatever the predictions are
  (right, wrong, random): predictions steer the work, never the result.
"""
import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "proto")
SRC, SO = os.path.join(HERE, "chain_group_model.c"), os.path.join(HERE, "libchaingroup.so")


def lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC, "-lm"])
    L = C.CDLL(SO)
    L.cgm_run.restype = C.c_longlong
    L.cgm_run.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                          C.c_double, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_longlong)]
    return L


def run(seed, D=12, div_mode=0, k=4.0, eps=0.05, delta=0.05, W0=2000.0, centre=1.0, sigma=0.03, far=0.3, groups=400, pred=2,
        contested=0.9):
    out = (C.c_longlong * 8)()
    bad = lib().cgm_run(seed, D, div_mode, k, eps, delta, W0, centre, sigma, far, groups, pred, contested, out)
    keys = ("fast_decided", "fast_undecided", "fast_wrong", "passes", "groups", "contested", "rejected")
    return bad, dict(zip(keys, list(out)))


# (the spread sigma puts the MC at its radius limit: with every dimension preferred r^2 ~ D sigma^2 / k against eps^2)
CASES = [
    dict(sigma=0.0305, far=0.3),                                     # C2-like: D = 12, k = 4 (a power of two: multiply by 1/k)
    dict(div_mode=1, k=3.0, sigma=0.0262, far=0.3),                  # k not a power of two: the reference's division
    dict(D=4, sigma=0.052, W0=300.0, far=0.3),                       # few markers, a light MC (coarse integer units)
    dict(D=15, sigma=0.0268, W0=50000.0, groups=200, far=0.3),       # widest record of the one-register layout, a heavy MC
    dict(delta=0.018, sigma=0.0185, far=0.3),                        # variances straddle delta^2: mixed preference masks
    dict(centre=40.0, eps=0.5, delta=0.5, sigma=0.3, far=0.3),       # far from the origin (cancellation in CF2 W - CF1^2)
    dict(far=0.0, contested=1.0, sigma=0.03), dict(far=1.0, contested=1.0, sigma=0.02), dict(contested=0.5, far=0.6),
]


@pytest.mark.parametrize("pred", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_group_schedules_reproduce_the_one_by_one_replay(case, pred):
    tot = {"fast_decided": 0, "fast_undecided": 0, "rejected": 0, "contested": 0, "passes": 0, "groups": 0}
    for seed in range(1, 7):
        bad, st = run(seed, pred=pred, **CASES[case])
        assert bad == 0, f"case {case} pred {pred} seed {seed}: {bad} groups differ from the one-by-one replay ({st})"
        assert st["fast_wrong"] == 0, f"case {case} seed {seed}: the fast test contradicted the exact test ({st})"
        for k in tot:
            tot[k] += st[k]
    # the regime is the intended one: both verdicts occur, and the fast test is a shortcut (it decides > 95 %)
    assert 0.03 * tot["contested"] < tot["rejected"] < 0.97 * tot["contested"]
    assert tot["fast_decided"] > 20 * max(tot["fast_undecided"], 1) or tot["fast_undecided"] == 0
    if pred == 3:  # right predictions: one pass per group (+ the exact first member, undecided cells)
        assert tot["passes"] <= 1.15 * tot["groups"] + 6
