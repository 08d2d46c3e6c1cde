"""ctypes wrapper around oracle/chronoclust_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by chronoclust_b200/.  Mirrors the reference's HDDStream contract
(/root/reference/chronoclust/clustering/hddstream.py:30-67, 89-128, 166-245) closely enough for
state-level comparisons: two ordered MC lists, id counters, final clusters.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64, _f64, _u8 = C.c_int64, C.c_double, C.c_uint8
_pi64, _pf64, _pu8 = C.POINTER(_i64), C.POINTER(_f64), C.POINTER(_u8)


def build(force=False):
    so = os.path.join(_HERE, "libcco.so")
    src = os.path.join(_HERE, "chronoclust_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libcco.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.cco_create.restype = C.c_void_p
        L.cco_create.argtypes = [C.c_int] + [_f64] * 8
        L.cco_destroy.argtypes = [C.c_void_p]
        L.cco_set_dnrm2.argtypes = [C.c_void_p, C.c_void_p]
        L.cco_begin_timepoint.argtypes = [C.c_void_p, _f64, _f64, _i64, C.c_int, _f64]
        L.cco_ingest.argtypes = [C.c_void_p, C.c_void_p, _i64, _i64, C.c_void_p, C.c_void_p]
        L.cco_offline.restype = _i64
        L.cco_offline.argtypes = [C.c_void_p]
        L.cco_count.restype = _i64
        L.cco_count.argtypes = [C.c_void_p, C.c_int]
        L.cco_dist_pairs.restype = _i64
        L.cco_dist_pairs.argtypes = [C.c_void_p]
        L.cco_last_id.restype = _i64
        L.cco_last_id.argtypes = [C.c_void_p, C.c_int]
        L.cco_export.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.cco_import_mc.argtypes = [C.c_void_p, C.c_int, _i64, _i64, _f64] + [C.c_void_p] * 4
        L.cco_set_counters.argtypes = [C.c_void_p, _i64, _i64]
        L.cco_set_thresholds.argtypes = [C.c_void_p, _f64, _f64, _i64]
        L.cco_n_clusters.restype = _i64
        L.cco_n_clusters.argtypes = [C.c_void_p]
        L.cco_n_members.restype = _i64
        L.cco_n_members.argtypes = [C.c_void_p]
        L.cco_export_clusters.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.cco_offline_m.restype = _i64
        L.cco_offline_m.argtypes = [C.c_void_p]
        L.cco_export_offline.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.cco_kat_projected_distance.restype = _f64
        L.cco_kat_projected_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.cco_kat_radius2.restype = _f64
        L.cco_kat_radius2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _f64, C.c_int]
        _LIB = L
    return _LIB


def scipy_dnrm2_pointer():
    """Address of the very dnrm2 numba binds for np.linalg.norm (scipy.linalg.cython_blas)."""
    import scipy.linalg.cython_blas as cb

    cap = cb.__pyx_capi__["dnrm2"]
    C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
    C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class MCList:
    """One exported list: parallel arrays in list order."""

    def __init__(self, ids, uids, w, cf1, cf2, cen, pref):
        self.ids, self.uids, self.w, self.cf1, self.cf2, self.cen, self.pref = ids, uids, w, cf1, cf2, cen, pref

    def __len__(self):
        return len(self.ids)


class OracleHDDStream:
    """Same constructor/entry point as the reference's HDDStream (hddstream.py:30, :166)."""

    def __init__(self, config, use_scipy_dnrm2=True):
        self.config = config
        self.epsilon = float(config["epsilon"])
        self.epsilon_squared = self.epsilon ** 2
        self.upsilon = float(config["upsilon"]) * self.epsilon
        self.delta = float(config["delta"])
        if self.delta > 1 or self.delta < 0:
            raise SystemExit("Given delta ({}) is out of range. Must be within 0-1.".format(self.delta))
        self.delta_squared = self.delta ** 2
        self.beta = float(config["beta"])
        self.k = float(config["k"])
        self.lambbda = float(config["lambda"])
        self.pi = self.mu = self.omicron = None
        self.last_data_timestamp = 0
        self.dataset_size = 0
        self.dataset_dimensionality = 0
        self._h = None
        self._use_scipy = use_scipy_dnrm2
        self.assign_uid = None
        self.stage = None

    def _ensure(self, D):
        if self._h is None:
            L = lib()
            self._h = L.cco_create(D, self.epsilon_squared, self.upsilon, self.upsilon ** 2, self.delta,
                                   self.delta_squared, self.beta, self.k, self.lambbda)
            if self._use_scipy:
                L.cco_set_dnrm2(self._h, scipy_dnrm2_pointer())
            self.dataset_dimensionality = D

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cco_destroy(self._h)
            self._h = None

    def set_dataset_dependent_parameters(self, X):
        # hddstream.py:89-128 (omicron from the PREVIOUS dataset size, mu from the current one)
        D = X.shape[1]
        cpi = float(self.config["pi"])
        self.pi = D if cpi <= 0 else round(cpi)
        self.omicron = self.config["omicron"] * self.dataset_size
        self.dataset_size = X.shape[0]
        self.mu = float(self.config["mu"]) * self.dataset_size

    def online_microcluster_maintenance(self, X, t, reset_param=True, offline=True):
        X = np.ascontiguousarray(X, dtype=np.float64)
        self._ensure(X.shape[1])
        if reset_param:
            self.set_dataset_dependent_parameters(X)
        L = lib()
        decay = (self.last_data_timestamp - t) != 0
        factor = 2 ** (-self.lambbda * (t - self.last_data_timestamp)) if decay else 1.0
        L.cco_begin_timepoint(self._h, float(self.mu), float(self.omicron), int(self.pi), int(decay), float(factor))
        N = X.shape[0]
        self.assign_uid = np.empty(N, np.int64)
        self.stage = np.empty(N, np.uint8)
        L.cco_ingest(self._h, _p(X), N, X.shape[1], _p(self.assign_uid), _p(self.stage))
        self.last_data_timestamp = t
        if offline:
            self.offline_clustering()

    def offline_clustering(self):
        return lib().cco_offline(self._h)

    def export(self, which):
        L = lib()
        n = L.cco_count(self._h, which)
        D = self.dataset_dimensionality
        ids, uids, w = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float64)
        arrs = [np.empty((n, D), np.float64) for _ in range(4)]
        L.cco_export(self._h, which, _p(ids), _p(uids), _p(w), *[_p(a) for a in arrs])
        return MCList(ids, uids, w, *arrs)

    def import_list(self, which, mcl):
        L = lib()
        for i in range(len(mcl)):
            L.cco_import_mc(self._h, which, int(mcl.ids[i]), int(mcl.uids[i]), float(mcl.w[i]),
                            *[_p(np.ascontiguousarray(a[i])) for a in (mcl.cf1, mcl.cf2, mcl.cen, mcl.pref)])

    @property
    def pcore(self):
        return self.export(0)

    @property
    def outlier(self):
        return self.export(1)

    @property
    def counters(self):
        L = lib()
        return L.cco_last_id(self._h, 0), L.cco_last_id(self._h, 1)

    @property
    def dist_pairs(self):
        return lib().cco_dist_pairs(self._h)

    def clusters(self):
        """[(member pcore ids in claim order, W, CF1, CF2, centroid, pref)] in emission order."""
        L = lib()
        nc, nm, D = L.cco_n_clusters(self._h), L.cco_n_members(self._h), self.dataset_dimensionality
        off, mem, w = np.zeros(nc + 1, np.int64), np.empty(nm, np.int64), np.empty(nc, np.float64)
        arrs = [np.empty((nc, D), np.float64) for _ in range(4)]
        L.cco_export_clusters(self._h, _p(off), _p(mem), _p(w), *[_p(a) for a in arrs])
        return [(mem[off[c]:off[c + 1]].tolist(), w[c], arrs[0][c], arrs[1][c], arrs[2][c], arrs[3][c])
                for c in range(nc)]

    def offline_intermediates(self):
        L = lib()
        M, D = L.cco_offline_m(self._h), self.dataset_dimensionality
        core, nbr, wn = np.zeros(M, np.uint8), np.zeros((M, M), np.uint8), np.zeros((M, M), np.uint8)
        w = np.zeros((M, D), np.float64)
        L.cco_export_offline(self._h, _p(core), _p(nbr), _p(wn), _p(w))
        return core, nbr, wn, w
