"""CPU suite: the block-speculative versioned commit (the algorithm of the CUDA engine, modelled with plain
loops in tests/proto/bsv_proto.c) reproduces the golden vectors of the live reference bit for bit, for any
block length / refinement depth / top-K / reject cap."""
import numpy as np
import pytest

from helpers import ORACLE_EXTRA_NAMES, STRESS_NAMES, assert_list_equal, config_of, load, stress_inputs
from proto.bsv import BsvHDDStream


def run_against(z, Xs, what, **kw):
    o = BsvHDDStream(config_of(z), **kw)
    for i, (t, X) in enumerate(zip(z["timestamps"].tolist(), Xs)):
        o.online_microcluster_maintenance(X, int(t))
        P = f"t{i}_"
        assert (o.assign_uid == z[P + "assign"]).all(), \
            f"{what} t{i}: assignment differs first at {np.flatnonzero(o.assign_uid != z[P + 'assign'])[:5]}"
        for which, name in ((0, "p_"), (1, "o_")):
            e = o.export(which)
            assert_list_equal((e.ids, e.uids, e.w, e.cf1, e.cf2, e.cen, e.pref), z, P + name, f"{what} t{i} list{which}")
        assert list(o.counters) == z[P + "counters"].tolist()
    return o


@pytest.mark.parametrize("kw", [dict(bmin=32, bmax=4096, itmax=4, topk=4, rmax=512),
                                dict(bmin=1, bmax=7, itmax=1, topk=1, rmax=3),
                                dict(bmin=512, bmax=512, itmax=8, topk=2, rmax=100000, contest=0.0),
                                dict(bmin=64, bmax=2048, itmax=3, topk=4, rmax=256, contest=1e9)])
@pytest.mark.parametrize("name", STRESS_NAMES + ORACLE_EXTRA_NAMES)
def test_bsv_model_matches_reference_on_stress(name, kw):
    z = load(f"stress_{name}.npz")
    o = run_against(z, stress_inputs(z), name, **kw)
    assert o.st.cells == sum(int(z[f"t{i}_assign"].shape[0]) for i in range(len(z["timestamps"])))


def test_bsv_model_matches_reference_on_c1():
    z = load("c1.npz")
    run_against(z, [z[f"scaled{t}"] for t in range(5)], "c1")
