"""Shared helpers of the parity tests: golden loading and state comparison (bit-exact)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STRESS_NAMES = ["downgrade", "feasibility", "churn", "tightdelta", "k1_gap", "c2like", "c3like"]
# golden vectors of the live reference in the saturated regime (config C5's smallest epsilon): they pin the oracle -- and the
# CPU model of the engine -- where the GPU suite compares the full-size C5 corner with the oracle (CPU suites only)
ORACLE_EXTRA_NAMES = ["saturated", "saturated_k3"]


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def stress_inputs(z):
    """Regenerates the inputs of a stress fixture from its seed and checks the stored checksum."""
    from chronoclust_b200.synth import gen

    ga = json.loads(str(z["gen"]))
    if "aniso" in ga and ga["aniso"] is not None:
        ga["aniso"] = tuple(ga["aniso"])
    Xs = gen(**ga)
    chk = np.array([float(x.sum()) for x in Xs])
    assert (chk == z["x_checksum"]).all(), "synthetic generator drifted from the golden inputs"
    return Xs


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    return a.shape == b.shape and (a.view(np.uint64) == b.view(np.uint64)).all()


def assert_list_equal(got, z, prefix, what):
    """got = (ids, uids, w, cf1, cf2, cen, pref) in list order; golden keys prefix+{ids,...}."""
    names = ["ids", "uids", "w", "cf1", "cf2", "cen", "pref"]
    exp = [z[prefix + n] for n in names]
    assert len(got[0]) == len(exp[0]), f"{what}: list length {len(got[0])} != {len(exp[0])}"
    for n, g, e in zip(names, got, exp):
        if n in ("ids", "uids"):
            assert (np.asarray(g) == e).all(), f"{what}: {n} differ"
        else:
            assert bits_equal(g, e), f"{what}: {n} not bit-identical (max abs diff {np.abs(np.asarray(g) - e).max()})"


def assert_clusters_equal(clusters, z, prefix, what):
    """clusters = [(member ids in claim order, w, cf1, cf2, cen, pref)] in emission order."""
    off = z[prefix + "cl_off"]
    assert len(clusters) == len(off) - 1, f"{what}: {len(clusters)} clusters != {len(off) - 1}"
    for c, (mem, w, cf1, cf2, cen, pref) in enumerate(clusters):
        s = set()
        for m in mem:
            s.add(int(m))
        exp_ids = z[prefix + "cl_idlist"][off[c]:off[c + 1]].tolist()
        assert list(s) == exp_ids, f"{what}: cluster {c} id set order {list(s)} != {exp_ids}"
        assert bits_equal(np.float64(w), z[prefix + "cl_w"][c]), f"{what}: cluster {c} weight"
        for n, g in (("cf1", cf1), ("cf2", cf2), ("cen", cen), ("pref", pref)):
            assert bits_equal(g, z[prefix + "cl_" + n][c]), f"{what}: cluster {c} {n} not bit-identical"


def config_of(z):
    return json.loads(str(z["config"]))
