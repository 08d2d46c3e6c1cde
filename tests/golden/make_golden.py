#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ by running the LIVE reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py [kat] [offline] [stress | stress:<name> ...] [c1]

What it pins:
  c1.npz            config C1 = the reference's own synthetic_dataset d0-d4 with the integration
                    test's parameters (chronoclust/tests/integration_test/normal_test.py:34-44):
                    raw + scaled inputs, and after every timepoint both MC lists (order, id,
                    prev_outlier_id, W, CF1, CF2, centroid, preference vector), the MC every point
                    went to, the id counters, and every offline cluster (CPython set iteration
                    order of its id, W, CF1, CF2, centroid, preference vector).
  c1_result.csv     result.csv written by the reference's app.run for those parameters + gating
                    file; checked here to be byte-identical to the reference's committed
                    expected_output/result.csv.
  c1_labels.npz     the cluster_id column of cluster_points_D0..4.csv from the same run (checked
                    against expected_output/ for identical ids and labels).
  stress_*.npz      generator-data runs that exercise the quirks (downgrades with skip-next rule,
                    outlier deletion, non-power-of-two decay, pi < D feasibility gate, k not a
                    power of two), and D=12 / D=40 runs shaped like configs C2 / C3.
  offline_sets.npz  randomised pcore-MC sets pushed through the reference's PreDeCon.run directly.
  kat.json          the known answers of the reference's unit tests, re-evaluated on the live code.
"""
import io
import json
import logging
import os
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

from chronoclust import app  # noqa: E402
from chronoclust.clustering.hddstream import HDDStream  # noqa: E402
from chronoclust.clustering.predecon import PreDeCon  # noqa: E402
from chronoclust.objects.predecon_mc import PredeconMC  # noqa: E402
from chronoclust.objects.microcluster import Microcluster  # noqa: E402
from chronoclust.scaling.scaler import Scaler  # noqa: E402
from chronoclust.utilities import mc_functions, predeconmc_functions  # noqa: E402
from chronoclust_b200.synth import gen  # noqa: E402

LOG = logging.getLogger("golden")
LOG.setLevel(logging.ERROR)


def dump_list(lst, D):
    n = len(lst)
    out = dict(ids=np.zeros(n, np.int64), uids=np.zeros(n, np.int64), w=np.zeros(n, np.float64),
               cf1=np.zeros((n, D)), cf2=np.zeros((n, D)), cen=np.zeros((n, D)), pref=np.zeros((n, D)))
    for i, m in enumerate(lst):
        out["ids"][i] = list(m.id)[0]
        out["uids"][i] = m.prev_outlier_id
        out["w"][i] = m.cumulative_weight
        out["cf1"][i], out["cf2"][i] = m.CF1, m.CF2
        out["cen"][i], out["pref"][i] = m.cluster_centroids, m.preferred_dimension_vector
    return out


def dump_state(h, N, D, prefix):
    out = {}
    for name, lst in (("p", h.pcore_MC), ("o", h.outlier_MC)):
        for k, v in dump_list(lst, D).items():
            out[f"{prefix}{name}_{k}"] = v
    assign = np.full(N, -1, np.int32)
    for m in list(h.pcore_MC) + list(h.outlier_MC):
        ks = np.fromiter(m.points.keys(), dtype=np.int64, count=len(m.points))
        assert (assign[ks] == -1).all()
        assign[ks] = m.prev_outlier_id
    assert (assign >= 0).all()
    out[f"{prefix}assign"] = assign
    out[f"{prefix}counters"] = np.array([h.pcore_MC_last_id, h.outlier_MC_last_id], np.int64)
    cl = h.final_clusters
    off = [0]
    mem = []
    for c in cl:
        mem.extend(list(c.id))  # CPython set iteration order
        off.append(len(mem))
    nc = len(cl)
    out[f"{prefix}cl_off"] = np.array(off, np.int64)
    out[f"{prefix}cl_idlist"] = np.array(mem, np.int64)
    out[f"{prefix}cl_w"] = np.array([float(c.cumulative_weight) for c in cl], np.float64)
    for k, attr in (("cf1", "CF1"), ("cf2", "CF2"), ("cen", "cluster_centroids"), ("pref", "preferred_dimension_vector")):
        out[f"{prefix}cl_{k}"] = np.array([np.asarray(getattr(c, attr), np.float64) for c in cl]).reshape(nc, D)
    return out


def run_reference(cfg, Xs, timestamps=None):
    h = HDDStream(dict(cfg), LOG)
    out = {}
    ts = timestamps or list(range(len(Xs)))
    for i, (t, X) in enumerate(zip(ts, Xs)):
        h.online_microcluster_maintenance(X, t)
        out.update(dump_state(h, X.shape[0], X.shape[1], f"t{i}_"))
    out["T"] = np.int64(len(Xs))
    out["timestamps"] = np.array(ts, np.int64)
    out["config"] = np.array(json.dumps(cfg))
    return out


def make_c1():
    data_dir = f"{REF}/chronoclust/tests/integration_test/test_files/dataset/full_dataset"
    files = [f"{data_dir}/synthetic_d{t}.csv.gz" for t in range(5)]
    cfg = {"beta": 0.2, "delta": 0.05, "epsilon": 0.03, "lambda": 2, "k": 4, "mu": 0.01, "pi": 3,
           "omicron": 0.000000435, "upsilon": 6.5}
    scaler = Scaler(files)
    raws = [pd.read_csv(f, header=0, sep=",").to_numpy() for f in files]
    Xs = [scaler.scale_data(r) for r in raws]
    out = run_reference(cfg, Xs)
    for t in range(5):
        out[f"raw{t}"] = raws[t]
        out[f"scaled{t}"] = np.ascontiguousarray(Xs[t])
    np.savez_compressed(f"{HERE}/c1.npz", **out)
    shutil.copy(f"{data_dir}/gating_centroids.csv", f"{HERE}/c1_gating_centroids.csv")

    # the real app.run, to pin the file outputs
    tmp = tempfile.mkdtemp()
    app.run(data=files, output_directory=tmp, gating_centroid_file=f"{data_dir}/gating_centroids.csv",
            param_beta=cfg["beta"], param_delta=cfg["delta"], param_epsilon=cfg["epsilon"],
            param_lambda=cfg["lambda"], param_k=cfg["k"], param_mu=cfg["mu"], param_pi=cfg["pi"],
            param_omicron=cfg["omicron"], param_upsilon=cfg["upsilon"])
    exp_dir = f"{REF}/chronoclust/tests/integration_test/test_files/expected_output"
    got, exp = open(f"{tmp}/result.csv", "rb").read(), open(f"{exp_dir}/result.csv", "rb").read()
    assert got == exp, "live reference result.csv differs from the committed expected_output"
    open(f"{HERE}/c1_result.csv", "wb").write(got)
    labels = {}
    for t in range(5):
        a = pd.read_csv(f"{tmp}/cluster_points_D{t}.csv", keep_default_na=False)
        b = pd.read_csv(f"{exp_dir}/cluster_points_D{t}.csv", keep_default_na=False)
        assert (a["id"].to_numpy() == b["id"].to_numpy()).all()
        assert (a["cluster_id"].astype(str).to_numpy() == b["cluster_id"].astype(str).to_numpy()).all()
        labels[f"ids{t}"] = a["id"].to_numpy().astype(np.int32)
        labels[f"labels{t}"] = a["cluster_id"].astype(str).to_numpy().astype("U")
        labels[f"xyz{t}"] = a[["x", "y", "z"]].to_numpy()
    np.savez_compressed(f"{HERE}/c1_labels.npz", **labels)
    logging.shutdown()
    shutil.rmtree(tmp, ignore_errors=True)
    print("c1 done")


STRESS = {
    # name: (gen args, config, timestamps)
    "downgrade": (dict(N=2500, D=3, T=4, C=4, seed=11),
                  {"beta": 0.9, "delta": 0.05, "epsilon": 0.06, "lambda": 1.5, "k": 4, "mu": 0.01, "pi": 3,
                   "omicron": 1e-3, "upsilon": 6.5}, None),
    "feasibility": (dict(N=2500, D=12, T=3, C=8, seed=12, aniso=(0.4, 3.0)),
                    {"beta": 0.2, "delta": 0.04, "epsilon": 0.12, "lambda": 0.5, "k": 3, "mu": 0.01, "pi": 8,
                     "omicron": 4.35e-6, "upsilon": 6.5}, None),
    "churn": (dict(N=3000, D=4, T=5, C=10, seed=15, drift=0.06),
              {"beta": 0.9, "delta": 0.05, "epsilon": 0.05, "lambda": 1.0, "k": 4, "mu": 0.01, "pi": 4,
               "omicron": 2e-3, "upsilon": 4.0}, None),
    "tightdelta": (dict(N=2500, D=6, T=3, C=6, seed=13),
                   {"beta": 0.2, "delta": 0.022, "epsilon": 0.04, "lambda": 1, "k": 7, "mu": 0.004, "pi": 4,
                    "omicron": 4.35e-6, "upsilon": 5.0}, None),
    "k1_gap": (dict(N=2000, D=5, T=3, C=5, seed=14),  # default k=1, timestamps with a gap and a repeat-free jump
               {"beta": 0.5, "delta": 0.0, "epsilon": 0.07, "lambda": 0.3, "k": 1, "mu": 0.005, "pi": 0,
                "omicron": 1e-4, "upsilon": 3.0}, [0, 2, 5]),
    "c2like": (dict(N=12000, D=12, T=2, C=20, seed=1234),
               {"beta": 0.2, "delta": 0.05, "epsilon": 0.05, "lambda": 2, "k": 4, "mu": 0.01, "pi": 12,
                "omicron": 4.35e-6, "upsilon": 6.5}, None),
    "c3like": (dict(N=4000, D=40, T=2, C=40, seed=1234),
               {"beta": 0.2, "delta": 0.05, "epsilon": 0.10, "lambda": 2, "k": 4, "mu": 0.01, "pi": 40,
                "omicron": 4.35e-6, "upsilon": 6.5}, None),
    # the saturated regime of the parameter sweep (config C5 at its smallest epsilon): the MCs sit at their radius limit,
    # clusters split over several pcore MCs, > 20 % of the cells reach the outlier stage.  They pin the ORACLE there (the
    # full-size C5 corner is compared with the oracle on the GPU); used by the CPU suites only (helpers.ORACLE_EXTRA_NAMES)
    "saturated": (dict(N=6000, D=12, T=2, C=20, seed=4321),
                  {"beta": 0.2, "delta": 0.05, "epsilon": 0.04, "lambda": 2, "k": 4, "mu": 0.01, "pi": 12,
                   "omicron": 4.35e-6, "upsilon": 6.5}, None),
    "saturated_k3": (dict(N=5000, D=6, T=3, C=6, seed=99),
                     {"beta": 0.8, "delta": 0.03, "epsilon": 0.03, "lambda": 1, "k": 3, "mu": 0.01, "pi": 6,
                      "omicron": 4.35e-6, "upsilon": 6.5}, None),
}


def make_stress(only=None):
    for name, (ga, cfg, ts) in STRESS.items():
        if only and name not in only:
            continue
        Xs = gen(**ga)
        out = run_reference(cfg, Xs, ts)
        out["gen"] = np.array(json.dumps(ga))
        # inputs are regenerated from the seed by the tests; keep a checksum to catch generator drift
        out["x_checksum"] = np.array([float(x.sum()) for x in Xs])
        np.savez_compressed(f"{HERE}/stress_{name}.npz", **out)
        print("stress", name, "done:", [int(out[f"t{i}_p_ids"].shape[0]) for i in range(len(Xs))],
              [int(out[f"t{i}_o_ids"].shape[0]) for i in range(len(Xs))],
              [int(out[f"t{i}_cl_w"].shape[0]) for i in range(len(Xs))])


def make_offline_sets():
    rng = np.random.default_rng(7)
    out = {}
    nset = 24
    for s in range(nset):
        D = int(rng.choice([3, 6, 12]))
        M = int(rng.integers(30, 120))
        k = float(rng.choice([3, 4, 15]))
        pi = int(rng.integers(D // 2, D + 1))
        delta = float(rng.choice([0.0005, 0.002, 0.01]))
        E = float(rng.choice([0.15, 0.25, 0.4]))
        nc = int(rng.integers(2, 7))
        ctr = rng.uniform(0.1, 0.9, size=(nc, D))
        sc = rng.uniform(0.01, 0.12, size=(nc, D))
        lab = rng.integers(0, nc, size=M)
        cen = ctr[lab] + rng.normal(0, 1, size=(M, D)) * sc[lab]
        w = rng.integers(1, 200, size=M).astype(np.float64) * rng.choice([1.0, 0.25, 0.7])
        cf1 = cen * w[:, None]
        cf2 = (cen ** 2 + rng.uniform(1e-4, 4e-3, size=(M, D))) * w[:, None]
        core = rng.random(M) < 0.6
        ids = rng.permutation(M * 2)[:M]
        dps = {}
        for i in range(M):
            dps[int(ids[i])] = PredeconMC(centroid=cen[i].copy(), id=int(ids[i]), is_core_cluster=bool(core[i]),
                                          cluster_CF1=cf1[i].copy(), cluster_CF2=cf2[i].copy(),
                                          cluster_cumulative_weight=float(w[i]))
        pre = PreDeCon(datapoints=dps, dataset_dimensionality=D, epsilon=E, delta=delta, lambbda=pi, mu=0.0, k=k)
        pre.run()
        P = f"s{s}_"
        out[P + "params"] = np.array([D, M, k, pi, delta, E], np.float64)
        out[P + "cen"], out[P + "w"], out[P + "cf1"], out[P + "cf2"] = cen, w, cf1, cf2
        out[P + "core"], out[P + "ids"] = core, ids.astype(np.int64)
        out[P + "subw"] = np.array([np.asarray(dps[int(i)].subspace_preference_vector, np.float64) for i in ids])
        nb = np.zeros((M, M), np.uint8)
        wn = np.zeros((M, M), np.uint8)
        pos = {int(v): i for i, v in enumerate(ids)}
        for i in range(M):
            for q in dps[int(ids[i])].neighbour_pts:
                nb[i, pos[q]] = 1
            for q in dps[int(ids[i])].weighted_neighbour_pts:
                wn[i, pos[q]] = 1
        out[P + "nbr"], out[P + "wnbr"] = np.packbits(nb), np.packbits(wn)
        off, mem = [0], []
        for c in pre.clusters:
            mem.extend(list(c.id))
            off.append(len(mem))
        ncl = len(pre.clusters)
        out[P + "cl_off"], out[P + "cl_idlist"] = np.array(off, np.int64), np.array(mem, np.int64)
        out[P + "cl_w"] = np.array([float(c.cumulative_weight) for c in pre.clusters])
        for kk, attr in (("cf1", "CF1"), ("cf2", "CF2"), ("cen", "cluster_centroids"),
                         ("pref", "preferred_dimension_vector")):
            out[P + "cl_" + kk] = np.array([np.asarray(getattr(c, attr), np.float64) for c in pre.clusters]).reshape(ncl, D)
    out["nset"] = np.int64(nset)
    np.savez_compressed(f"{HERE}/offline_sets.npz", **out)
    print("offline sets done")


def make_kat():
    """Known answers of the reference's unit tests: the inputs are the vectors of
    chronoclust/tests/objects_test/unittest_microcluster.py and clustering_test/unittest_predecon.py,
    the outputs are re-evaluated on the live numba functions (full precision, not the rounded asserts)."""
    kat = {}
    # unittest_microcluster.py:10-32 (projected distance 0.85 / 1.25 to 2 dp)
    kat["projdist"] = []
    for cen, expect in (([0.1, 0.2, 0.03], 0.85), ([-0.1, 0.2, -0.03], 1.25)):
        m = Microcluster(cf1=np.zeros(3), cf2=np.zeros(3), cluster_centroids=cen,
                         preferred_dimension_vector=[1.0, 15.0, 15.0])
        d = float(m.get_projected_dist_to_point([1.0, 0.5, 0.7]))
        assert round(d, 2) == expect
        kat["projdist"].append(dict(cen=cen, pref=[1.0, 15.0, 15.0], pt=[1.0, 0.5, 0.7], dist=d, rounded=expect))
    # unittest_microcluster.py:34-80 (preference vector under delta^2 = 0.01 / 0.05 / 0.1, k = 15)
    points = [[0.17550518, 0.50150137, 0.0715026, 0.46715915, 0.11825116],
              [0.09084978, 0.33935363, 0.06932869, 0.78185322, 0.62759489],
              [0.22507306, 0.02771729, 0.46630673, 0.75367467, 0.2201496],
              [0.26507548, 0.44774516, 0.28568398, 0.80777178, 0.12095075],
              [0.43343372, 0.35738624, 0.4001447, 0.89195078, 0.29652304],
              [0.48627326, 0.52784397, 0.22927219, 0.801923, 0.07897944],
              [0.31972963, 0.29667314, 0.20070554, 0.31300255, 0.4958211],
              [0.05191981, 0.76440696, 0.0478006, 0.0201296, 0.25368318],
              [0.18290483, 0.65387882, 0.174167, 0.21822311, 0.2230557],
              [0.87574659, 0.77501901, 0.21127804, 0.15939672, 0.6381301]]
    kat["prefvec"] = dict(pts=points, cases=[])
    for d2, expect in ((0.01, [1, 1, 1, 1, 1]), (0.05, [1, 15, 15, 1, 15]), (0.1, [15, 15, 15, 15, 15])):
        mc = Microcluster(cf1=np.zeros(5), cf2=np.zeros(5))
        for idx, pnt in enumerate(points):
            mc.add_new_point(np.array(pnt), 0, idx)
            mc.update_preferred_dimensions(d2, 15)
        assert np.asarray(mc.preferred_dimension_vector).tolist() == expect
        kat["prefvec"]["cases"].append(dict(delta2=d2, k=15.0, pref=[float(v) for v in mc.preferred_dimension_vector],
                                            cf1=mc.CF1.tolist(), cf2=mc.CF2.tolist(),
                                            cen=np.asarray(mc.cluster_centroids).tolist(), w=float(mc.cumulative_weight)))
    # unittest_microcluster.py:82-104 (radius^2 = 0.1551429607662637 to 10 places)
    cf1 = [0.68756544, 0.96853843, 0.41156436, 0.13236377, 0.12836222, 0.55662013, 0.9671396, 0.99469293, 0.86402299,
           0.90838236, 0.52934492, 0.37423623, 0.02787237, 0.35216188, 0.96222637, 0.09291304, 0.08972414, 0.76429683,
           0.78941125, 0.53722776]
    cf2 = [4.72746229e-01, 9.38066699e-01, 1.69385220e-01, 1.75201686e-02, 1.64768583e-02, 3.09825969e-01,
           9.35359004e-01, 9.89414034e-01, 7.46535721e-01, 8.25158518e-01, 2.80206042e-01, 1.40052759e-01,
           7.76869185e-04, 1.24017991e-01, 9.25879595e-01, 8.63283346e-03, 8.05042136e-03, 5.84149644e-01,
           6.23170114e-01, 2.88613669e-01]
    pref = [1, 1, 16, 16, 1, 16, 16, 16, 16, 16, 1, 16, 1, 16, 16, 16, 1, 1, 1, 16]
    mc = Microcluster(cf1=np.array(cf1), cf2=np.array(cf2), preferred_dimension_vector=np.array(pref),
                      cumulative_weight=20)
    r2 = float(mc.calculate_projected_radius_squared())
    assert abs(r2 - 0.1551429607662637) < 1e-10
    kat["radius2"] = dict(cf1=cf1, cf2=cf2, pref=[float(v) for v in pref], w=20.0, r2=r2)
    # unittest_predecon.py:8-15 (Euclidean 5.196152422706632), :41-47 (weighted dist^2 37.19)
    a, b = np.array([1, 5, 6, 3, 2], dtype='float'), np.array([6, 4, 6, 4, 2], dtype='float')
    kat["euclid"] = dict(a=a.tolist(), b=b.tolist(), dist=float(predeconmc_functions.calculate_euclidean_dist(a, b)))
    pv, p_, q_ = np.array([15.0, 1, 1, 15]), np.array([0.1, 4.5, 4.2, 3.0]), np.array([1.1, 4.3, 2.2, 4.1])
    kat["wdist2"] = dict(pref=pv.tolist(), p=p_.tolist(), q=q_.tolist(),
                         dist=float(predeconmc_functions.calculate_weighted_dist_squared(pv, p_, q_)))
    # unittest_predecon.py:17-39 (variance along a dimension, 3 dp)
    point = [0.187, 0.922, 0.896, 0.098, 0.707, 0.626, 0.447, 0.588, 0.752, 0.041]
    neighbours = [
        [0.873, 0.179, 0.585, 0.036, 0.051, 0.708, 0.485, 0.75, 0.665, 0.019],
        [0.218, 0.791, 0.451, 0.061, 0.197, 0.083, 0.453, 0.538, 0.136, 0.046],
        [0.314, 0.119, 0.153, 0.336, 0.174, 0.125, 0.02, 0.752, 0.89, 0.147],
        [0.21, 0.681, 0.018, 0.503, 0.081, 0.612, 0.395, 0.458, 0.071, 0.992],
        [0.26, 0.59, 0.788, 0.063, 0.466, 0.702, 0.387, 0.204, 0.91, 0.888],
        [0.775, 0.173, 0.92, 0.854, 0.034, 0.511, 0.933, 0.237, 0.375, 0.891],
        [0.441, 0.021, 0.142, 0.754, 0.121, 0.626, 0.661, 0.618, 0.967, 0.345],
        [0.457, 0.708, 0.322, 0.715, 0.075, 0.212, 0.481, 0.347, 0.935, 0.234],
        [0.516, 0.052, 0.745, 0.137, 0.764, 0.515, 0.888, 0.948, 0.362, 0.912],
        [0.287, 0.385, 0.658, 0.735, 0.354, 0.317, 0.321, 0.995, 0.071, 0.864]]
    var = [float(predeconmc_functions.calculate_variance_along_dimension(np.array(point[i]),
                                                                          np.array([n[i] for n in neighbours])))
           for i in range(10)]
    assert np.round(var, 3).tolist() == [0.109, 0.385, 0.261, 0.202, 0.275, 0.085, 0.068, 0.07, 0.173, 0.392]
    kat["nbrvar"] = dict(point=point, neighbours=neighbours, var=var)
    json.dump(kat, open(f"{HERE}/kat.json", "w"), indent=1)
    print("kat done")


if __name__ == "__main__":
    which = sys.argv[1:] or ["kat", "offline", "stress", "c1"]
    if "kat" in which:
        make_kat()
    if "offline" in which:
        make_offline_sets()
    if "stress" in which:
        make_stress()
    only = [w.split(":", 1)[1] for w in which if w.startswith("stress:")]  # e.g. stress:saturated
    if only:
        make_stress(only)
    if "c1" in which:
        make_c1()
