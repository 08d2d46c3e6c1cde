"""ctypes driver of tests/proto/bsv_proto.c (CPU model of the block-speculative versioned commit).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("blocks", "iters", "cells", "mismatches", "unknown_cuts", "iter_cuts",
                                         "upgrades", "max_iters", "rejects", "tk_miss")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def lib():
    global _LIB
    if _LIB is None:
        so, src = os.path.join(_HERE, "libbsvproto.so"), os.path.join(_HERE, "bsv_proto.c")
        dep = os.path.join(_HERE, "..", "..", "oracle", "chronoclust_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-std=c11",
                                   "-o", so, src, "-lm"])
        L = C.CDLL(so)
        L.cco_create.restype = C.c_void_p
        L.cco_create.argtypes = [C.c_int] + [C.c_double] * 8
        L.cco_destroy.argtypes = [C.c_void_p]
        L.cco_set_dnrm2.argtypes = [C.c_void_p, C.c_void_p]
        L.cco_begin_timepoint.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_int, C.c_double]
        L.cco_ingest_bsv.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_void_p]
        L.cco_offline.restype = C.c_int64
        L.cco_offline.argtypes = [C.c_void_p]
        L.cco_count.restype = C.c_int64
        L.cco_count.argtypes = [C.c_void_p, C.c_int]
        L.cco_last_id.restype = C.c_int64
        L.cco_last_id.argtypes = [C.c_void_p, C.c_int]
        L.cco_export.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        _LIB = L
    return _LIB


class BsvHDDStream(_o.OracleHDDStream):
    """OracleHDDStream whose ordered loop is replaced by the block-speculative model."""

    def __init__(self, config, bmin=32, bmax=4096, itmax=4, topk=4, rmax=512, contest=4.0):
        super().__init__(config, use_scipy_dnrm2=False)
        self.bp = (bmin, bmax, itmax, topk, rmax, contest)
        self.st = Stats()

    def _ensure(self, D):
        if self._h is None:
            self._h = lib().cco_create(D, self.epsilon_squared, self.upsilon, self.upsilon ** 2, self.delta,
                                       self.delta_squared, self.beta, self.k, self.lambbda)
            self.dataset_dimensionality = D

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cco_destroy(self._h)
            self._h = None

    def online_microcluster_maintenance(self, X, t, reset_param=True, offline=False):
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
        bmin, bmax, itmax, topk, rmax, contest = self.bp
        L.cco_ingest_bsv(self._h, _o._p(X), N, X.shape[1], _o._p(self.assign_uid), _o._p(self.stage), bmin, bmax, itmax,
                         topk, rmax, contest, C.byref(self.st))
        self.last_data_timestamp = t

    def export(self, which):
        L = lib()
        n = L.cco_count(self._h, which)
        D = self.dataset_dimensionality
        ids, uids, w = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float64)
        arrs = [np.empty((n, D), np.float64) for _ in range(4)]
        L.cco_export(self._h, which, _o._p(ids), _o._p(uids), _o._p(w), *[_o._p(a) for a in arrs])
        return _o.MCList(ids, uids, w, *arrs)

    @property
    def counters(self):
        L = lib()
        return L.cco_last_id(self._h, 0), L.cco_last_id(self._h, 1)
