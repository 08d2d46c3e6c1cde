"""world_size-2 gloo test (CPU) of the row-sharded offline path's host logic: row partitioning, padding, the all-gathers,
the CSR exchange (per-rank lists merged by an all-reduce) and the replicated cluster stage.  The compute stages are
injected: here they are backed by the oracle's full-matrix intermediates (test infrastructure), on the GPU by the C-ABI
kernels (tests/test_gpu_parity.py::test_sharded_offline_nccl_matches_single_rank runs those on 2 GPUs)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleStages:
    """CPU stand-in for CudaStages: slices rows out of the oracle's intermediates."""

    def __init__(self, core, nbr, wn, subw, k, labels):
        self.torch = torch
        self.core, self.nbr, self.wn, self.subw, self.k, self.labels = core, nbr, wn, subw, k, labels

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    @staticmethod
    def _pack(rows_bits, words):
        out = np.zeros((rows_bits.shape[0], words), np.uint32)
        for q in range(rows_bits.shape[1]):
            out[:, q >> 5] |= (rows_bits[:, q].astype(np.uint32) << np.uint32(q & 31))
        return out.view(np.int32)

    def neighbours(self, cen, M, D, r0, r1, E, E2, nbr, cnt):
        words = (M + 31) // 32
        nbr[:r1 - r0] = torch.from_numpy(self._pack(self.nbr[r0:r1], words))
        cnt[:r1 - r0] = torch.from_numpy(self.nbr[r0:r1].sum(1).astype(np.int32))
        return 0

    def subspace(self, cen, M, D, r0, r1, nbr, cnt, delta, submask):
        m = np.zeros(r1 - r0, np.int64)
        for i in range(r0, r1):
            m[i - r0] = sum(1 << d for d in range(D) if self.subw[i, d] == self.k and self.k != 1.0)
        submask[:r1 - r0] = torch.from_numpy(m)

    def weighted(self, cen, M, D, r0, r1, nbr, submask_all, k, E2, wnbr):
        assert submask_all.shape[0] >= M
        words = (M + 31) // 32
        wnbr[:r1 - r0] = torch.from_numpy(self._pack(self.wn[r0:r1], words))

    def rowinfo(self, wn_rows, M, r0, r1, iso, nnz):
        rows = self.wn[r0:r1].astype(bool)
        c = rows.sum(1)
        lone = (c == 1) & rows[np.arange(r1 - r0), np.arange(r0, r1)]
        iso[:r1 - r0] = torch.from_numpy(lone.astype(np.uint8))
        nnz[:r1 - r0] = torch.from_numpy(np.where(lone, 0, c).astype(np.int32))

    def fill(self, wn_rows, M, r0, r1, iso_all, off_all, col):
        for row in range(r0, r1):
            if not int(iso_all[row]):
                cols = np.flatnonzero(self.wn[row]).astype(np.int32)
                o = int(off_all[row])
                assert int(off_all[row + 1]) - o == len(cols)
                col[o:o + len(cols)] = torch.from_numpy(cols)

    def clusters_csr(self, M, off_all, col, iso_all, core, submask_all, k, pi):
        # the merged CSR must be the CSR of the oracle's full matrix without its isolated rows
        full = self.wn.astype(bool)
        lone = (full.sum(1) == 1) & full[np.arange(M), np.arange(M)]
        assert (iso_all[:M].numpy().astype(bool) == lone).all(), "gathered isolated flags differ"
        exp = np.concatenate([np.flatnonzero(full[r]) for r in range(M) if not lone[r]] + [np.zeros(0, np.int64)])
        assert int(off_all[M]) == len(exp) and (col[:len(exp)].numpy() == exp).all(), "merged CSR differs"
        self.used_csr = True
        return self.labels, np.arange(M, dtype=np.int32), np.zeros(1, np.int32), 0

    def clusters(self, M, wnbr_all, core, submask_all, k, pi):
        # the gathered matrix must equal the oracle's full one: that is what this test is about
        words = (M + 31) // 32
        full = self._pack(self.wn, words)
        assert (wnbr_all[:M].numpy() == full).all(), "gathered WN rows differ from the full matrix"
        return self.labels, np.arange(M, dtype=np.int32), np.zeros(1, np.int32), 0


class _NoCsr:
    """View of a stage object without the CSR methods (forces the bit-row all-gather)."""

    def __init__(self, st):
        self._st = st
        self.torch = st.torch

    def __getattr__(self, name):
        if name in ("rowinfo", "fill", "clusters_csr"):
            raise AttributeError(name)
        return getattr(self._st, name)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import load
        from chronoclust_b200.offline_sharded import row_range, sharded_offline
        from oracle.oracle import OracleHDDStream, lib as olib, _p

        z = load("offline_sets.npz")
        ok = True
        for s in (0, 3, 5):
            P = f"s{s}_"
            D, M, k, pi, delta, E = z[P + "params"]
            D, M, pi = int(D), int(M), int(pi)
            cen, w, cf1, cf2, ids = (z[P + n] for n in ("cen", "w", "cf1", "cf2", "ids"))
            cfg = {"beta": 0.0, "delta": float(delta), "epsilon": 1e150, "lambda": 0, "k": float(k), "mu": 0.0, "pi": pi,
                   "omicron": 0.0, "upsilon": float(E) / 1e150}
            o = OracleHDDStream(cfg)
            o._ensure(D)
            for i in range(M):
                olib().cco_import_mc(o._h, 0, int(ids[i]), int(ids[i]), float(w[i]), _p(np.ascontiguousarray(cf1[i])),
                                     _p(np.ascontiguousarray(cf2[i])), _p(np.ascontiguousarray(cen[i])), _p(np.ones(D)))
            olib().cco_set_thresholds(o._h, 0.0, 0.0, pi)
            o.offline_clustering()
            core, nbr, wn, subw = o.offline_intermediates()
            labels = np.arange(M, dtype=np.int32)
            st = OracleStages(core, nbr, wn, subw, float(k), labels)
            lab, order, cl_off, ncl, info = sharded_offline(st, torch.from_numpy(cen), torch.from_numpy(core), M, D,
                                                            float(k), pi, float(delta), float(E), float(E) ** 2,
                                                            group=None, dist=dist)
            R, r0, r1 = row_range(M, world, rank)
            ok = ok and info["rows"] == (r0, r1) and (lab == labels).all()
            ok = ok and info["neighbour_count"] == int(nbr[r0:r1].sum())
            ok = ok and info["exchange"] in ("csr", "bitrows") and (info["exchange"] == "bitrows" or st.used_csr)
            # the same through the bit-row exchange (what a stage object without the CSR methods gets)
            st2 = OracleStages(core, nbr, wn, subw, float(k), labels)
            lab2, *_rest, info2 = sharded_offline(_NoCsr(st2), torch.from_numpy(cen), torch.from_numpy(core), M, D, float(k), pi,
                                                  float(delta), float(E), float(E) ** 2, group=None, dist=dist)
            ok = ok and info2["exchange"] == "bitrows" and (lab2 == labels).all()
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_sharded_offline_gloo_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()  # no fork() of the (multi-threaded) pytest process
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_row_range_covers_everything():
    from chronoclust_b200.offline_sharded import row_range

    for M in (0, 1, 5, 31, 32, 33, 1000, 100000):
        for world in (1, 2, 4, 8):
            seen = []
            for rank in range(world):
                R, r0, r1 = row_range(M, world, rank)
                assert 0 <= r0 <= r1 <= M and r1 - r0 <= R
                seen.extend(range(r0, r1))
            assert seen == list(range(M))
