"""GPU localizer (Localizer::Compact on the device) and the raw-id fused step: bit-exact integer work."""
import numpy as np
import pytest

from conftest import parse_kwargs, syn_batches
from oracle import oracle as O
from util import assert_close, localized, oracle_state, rand_batch

pytestmark = pytest.mark.gpu
capi = pytest.importorskip("difacto_b200.capi")


def test_localizer_goldens_on_gpu(rcv1, refout):
    E = capi.Engine(V_dim=0, table_capacity=1 << 12)
    lidx, keys, cnt = E.localize(rcv1["offset"], rcv1["index"])
    assert np.array_equal(lidx, refout["loc_lidx"])           # output of the compiled reference
    assert np.array_equal(keys, refout["loc_keys"])
    assert np.array_equal(cnt, refout["loc_cnt"])
    assert int(O.reverse_bytes_np(keys).sum()) == 65111856 and cnt.sum() == 9648     # localizer_test.cc:26-27
    l2, k2, c2 = E.localize(rcv1["offset"], rcv1["index"], max_index=1000)
    assert np.array_equal(l2, refout["loc1000_lidx"]) and np.array_equal(k2, refout["loc1000_keys"])
    assert np.array_equal(c2, refout["loc1000_cnt"])
    assert int(O.reverse_bytes_np(k2).sum()) == 478817                                 # localizer_test.cc:48


@pytest.mark.parametrize("case", ["small_ids", "full_64bit", "all_same", "max_id", "empty_rows"])
def test_localizer_edge_cases_bit_exact(case):
    rng = np.random.default_rng(5)
    E = capi.Engine(V_dim=0, table_capacity=1 << 12)
    B = 300
    nnzr = rng.integers(0, 50, B)
    if case == "empty_rows":
        nnzr[::2] = 0
    off = np.concatenate([[0], np.cumsum(nnzr)]).astype(np.uint64)
    n = int(off[-1])
    if case == "small_ids":
        idx = rng.integers(0, 1000, n).astype(np.uint64)
    elif case == "full_64bit":
        idx = rng.integers(0, 2 ** 63, n).astype(np.uint64) * np.uint64(2) + rng.integers(0, 2, n).astype(np.uint64)
    elif case == "all_same":
        idx = np.full(n, 123456789, np.uint64)
    elif case == "max_id":
        idx = rng.integers(0, 50, n).astype(np.uint64)
        idx[::7] = np.uint64(0xFFFFFFFFFFFFFFFF)     # id % (2^64-1) == 0 (localizer.cc:24)
        idx[1::7] = np.uint64(0)
    else:
        idx = rng.integers(0, 10 ** 9, n).astype(np.uint64)
    ol, ok, oc = O.localize(off, idx)
    gl, gk, gc = E.localize(off, idx)
    assert np.array_equal(gl, ol) and np.array_equal(gk, ok) and np.array_equal(gc, oc)
    assert np.all(gk[1:] > gk[:-1])


def test_localizer_zero_rows():
    E = capi.Engine(V_dim=0, table_capacity=1 << 10)
    l, k, c = E.localize(np.zeros(1, np.uint64), np.zeros(0, np.uint64))
    assert len(l) == 0 and len(k) == 0
    l, k, c = E.localize(np.zeros(4, np.uint64), np.zeros(0, np.uint64))
    assert len(l) == 0 and len(k) == 0


@pytest.mark.parametrize("V_dim,valued,scatter", [(16, True, "sorted"), (64, False, "sorted"), (16, False, "atomic"),
                                                  (5, True, "sorted"), (0, True, "sorted")])
def test_raw_step_equals_localized_step_and_oracle(V_dim, valued, scatter):
    rng = np.random.default_rng(77 + V_dim)
    kw = dict(V_dim=V_dim, l1=0.2, l2=0.01, lr=0.2, V_lr=0.05, V_threshold=3, V_l2=0.02, V_init_scale=0.2, seed=5)
    batches = [rand_batch(rng, 150, 30, 400, valued and j % 2 == 0) for j in range(4)]
    M = O.Oracle(**kw)
    exact = scatter == "sorted" and V_dim in (8, 16, 32, 64, 128)    # the atomic-free path is bit-reproducible
    R = capi.Engine(table_capacity=1 << 14, scatter=scatter, **kw)     # raw ids -> GPU localizer
    L = capi.Engine(table_capacity=1 << 14, scatter=scatter, **kw)     # host-localized
    for ep in range(3):
        for (o, l, i, v) in batches:
            b = localized((o, l, i, v))
            ref = M.sgd_step(o, i, v, l, True, ep == 0)
            pr = R.train_step_raw(o, i, v, l, push_cnt=(ep == 0), is_train=True)
            pl = L.train_step(o, b["lidx"], v, l, b["keys"], b["cnt"] if ep == 0 else None, True)
            assert pr.nrows == ref[4]
            assert abs(pr.loss - ref[0]) <= 1e-4 * abs(ref[0]) + 1e-4
            assert abs(pr.penalty - ref[1]) <= 1e-4 * abs(ref[1]) + 1e-5
            if exact:
                assert pr.loss == pl.loss and pr.penalty == pl.penalty     # same kernels, same CSC order
    keys = np.unique(np.concatenate([O.reverse_bytes_np(b[2]) for b in batches]))
    sr, sl = R.read_entries(keys), L.read_entries(keys)
    if exact:
        for a, b in zip(sr, sl):
            assert np.array_equal(a, b)        # bit-identical to the host-localized path
    oscal, ohasv, oV, ocg = oracle_state(M, keys)
    assert np.array_equal(sr[1], ohasv)
    assert_close(sr[0][:, 1:], oscal[:, 1:], what="w/sqrt_g/z", rtol=1e-3, atol=1e-5)
    assert_close(sr[2], oV, what="V", rtol=1e-3, atol=1e-5)
    assert R.rng_state() == M.seed()


def test_raw_async_pipeline():
    rng = np.random.default_rng(3)
    kw = dict(V_dim=32, l1=0.05, lr=0.1, V_threshold=1, seed=4)
    batches = [rand_batch(rng, 200, 25, 800, False) for _ in range(6)]
    A, S = capi.Engine(table_capacity=1 << 14, **kw), capi.Engine(table_capacity=1 << 14, **kw)
    tot = 0.0
    for (o, l, i, v) in batches:
        tot += S.train_step_raw(o, i, v, l, push_cnt=True).loss
    for j, (o, l, i, v) in enumerate(batches):
        A.train_step_raw_async(len(l), o, i, None, l, push_cnt=True)
        if j + 1 < len(batches) and j % 2 == 0:      # prefetch every other batch: both paths interleave
            o2, l2, i2, _ = batches[j + 1]
            A.prefetch_raw(len(l2), o2, i2, None, l2)
    got = sum(A.wait_step().loss for _ in batches)
    assert abs(got - tot) <= 1e-5 * abs(tot)
    keys = np.unique(np.concatenate([O.reverse_bytes_np(b[2]) for b in batches]))
    for a, b in zip(A.read_entries(keys), S.read_entries(keys)):
        assert np.array_equal(a, b)


def test_full_size_localizer_properties():
    B, NNZ = 65536, 100
    rng = np.random.default_rng(9)
    ids = rng.integers(0, 10 ** 9, B * NNZ).astype(np.uint64)
    off = (np.arange(B + 1, dtype=np.uint64) * np.uint64(NNZ))
    E = capi.Engine(V_dim=0, table_capacity=1 << 12)
    lidx, keys, cnt = E.localize(off, ids)
    rk = O.reverse_bytes_np(ids)
    ukeys, inv, ucnt = np.unique(rk, return_inverse=True, return_counts=True)
    assert np.array_equal(keys, ukeys)                      # sorted unique reversed keys
    assert np.array_equal(lidx, inv.astype(np.uint32))      # rank of every nnz
    assert np.array_equal(cnt, ucnt.astype(np.float32))
    assert np.array_equal(keys[lidx], rk)                   # round trip: the remap is the inverse of unique
