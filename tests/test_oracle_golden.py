"""Pin the C oracle (oracle/fm_oracle.c) against the reference's own golden values and
against outputs of the compiled reference committed under tests/golden/.

Golden sources: tests/cpp/fm_loss_test.cc:35,39,78,82; tests/cpp/localizer_test.cc:26-27,
48-49,56-63; tests/cpp/sgd_learner_test.cc:10-30,45 (all in /root/reference).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import parse_kwargs, syn_batches
from oracle import oracle as O

SGD_GOLDEN = [69.314718, 69.314718, 67.151912, 61.414778, 56.244989, 53.218700, 51.248737,
              49.846688, 48.650164, 47.698351, 46.924038, 46.388223, 45.970721, 45.499307,
              45.102245, 44.798413, 44.565211, 44.386417, 44.240657, 44.109764]


def test_reverse_bytes_involution():
    # localizer_test.cc:56-63
    mx = 0xFFFFFFFFFFFFFFFF
    n = 1000000
    j = (np.arange(0, n, 997, dtype=np.uint64) * np.uint64(mx // n))
    assert np.array_equal(O.reverse_bytes_np(O.reverse_bytes_np(j)), j)
    for v in [0, 1, 0x123456789ABCDEF0, mx, 47149]:
        assert int(O.reverse_bytes(v)) == int(O.reverse_bytes_np(np.uint64(v)))
        assert int(O.reverse_bytes(O.reverse_bytes(v))) == v
    assert int(O.reverse_bytes(0x1)) == 0x1000000000000000
    assert int(O.reverse_bytes(0x123456789ABCDEF0)) == 0x0FEDCBA987654321


def test_rand_r_matches_glibc():
    libc = C.CDLL("libc.so.6")
    libc.rand_r.argtypes = [C.POINTER(C.c_uint)]
    for seed in (0, 1, 12345, 0xFFFFFFFF):
        a, b = C.c_uint(seed), C.c_uint(seed)
        for _ in range(20000):
            assert O.orc().orc_rand_r(C.byref(a)) == libc.rand_r(C.byref(b))
        assert a.value == b.value


def test_localizer_goldens(rcv1, refout):
    lidx, keys, cnt = O.localize(rcv1["offset"], rcv1["index"])
    uidx = O.reverse_bytes_np(keys)
    assert int(uidx.sum()) == 65111856          # localizer_test.cc:26
    assert float(cnt.sum()) == 9648.0           # localizer_test.cc:27
    assert np.all(np.diff(keys.astype(np.float64)) > 0) and len(np.unique(keys)) == len(keys)
    assert np.array_equal(lidx, refout["loc_lidx"])
    assert np.array_equal(keys, refout["loc_keys"])
    assert np.array_equal(cnt, refout["loc_cnt"])
    l2, k2, c2 = O.localize(rcv1["offset"], rcv1["index"], max_index=1000)
    assert int(O.reverse_bytes_np(k2).sum()) == 478817   # localizer_test.cc:48
    assert float(c2.sum()) == 9648.0
    assert np.array_equal(l2, refout["loc1000_lidx"]) and np.array_equal(k2, refout["loc1000_keys"])


def test_fm_loss_nov(rcv1, refout):
    lidx, keys, _ = O.localize(rcv1["offset"], rcv1["index"])
    uidx = O.reverse_bytes_np(keys)
    w = (uidx.astype(np.float64) / 5e4).astype(np.float32)
    pred = O.fm_predict(0, rcv1["offset"], lidx, rcv1["value"], w)
    objv = O.evaluate(rcv1["label"], pred)
    assert abs(objv - 147.4672) < 1e-3           # fm_loss_test.cc:35
    g = O.fm_calc_grad(0, rcv1["offset"], lidx, rcv1["value"], rcv1["label"], w, pred)
    assert abs(float((g.astype(np.float64) ** 2).sum()) - 90.5817) < 1e-3   # :39
    assert np.array_equal(pred, refout["nov_pred"])
    assert np.array_equal(g, refout["nov_grad"])


def test_fm_loss_hasv(rcv1, refout):
    lidx, keys, _ = O.localize(rcv1["offset"], rcv1["index"])
    U, k = len(keys), 5
    W = refout["hasv_w"]
    w_pos = (np.arange(U) * (k + 1)).astype(np.int32)
    V_pos = w_pos + 1
    pred = O.fm_predict(k, rcv1["offset"], lidx, rcv1["value"], W, w_pos, V_pos)
    assert abs(O.evaluate(rcv1["label"], pred) - 330.628) < 1e-3           # fm_loss_test.cc:78
    g = O.fm_calc_grad(k, rcv1["offset"], lidx, rcv1["value"], rcv1["label"], W, pred, w_pos, V_pos)
    assert abs(float((g.astype(np.float64) ** 2).sum()) - 1.2378e3) < 1e-1  # :82
    assert np.array_equal(pred, refout["hasv_pred"])
    assert np.array_equal(g, refout["hasv_grad"])


def test_sgd_learner_golden_trace(rcv1, refout):
    # sgd_learner_test.cc:9-49 with stop_rel_objv=0 so that all 20 epochs run
    M = O.Oracle(V_dim=0, l1=1, l2=1, lr=1)
    for ep in range(20):
        pr = M.sgd_step(rcv1["offset"], rcv1["index"], rcv1["value"], rcv1["label"], True, ep == 0)
        assert abs(float(pr[0]) - SGD_GOLDEN[ep]) < 5e-5
        assert abs(float(pr[0]) - float(refout["sgd_v0_trace"][ep, 0])) < 5e-5


def test_sgd_v8_trace_vs_reference(rcv1, refout):
    kw = parse_kwargs(refout["sgd_v8_kwargs"])
    M = O.Oracle(**kw)
    lidx, keys, _ = O.localize(rcv1["offset"], rcv1["index"])
    for ep in range(len(refout["sgd_v8_trace"])):
        pr = M.sgd_step(rcv1["offset"], rcv1["index"], rcv1["value"], rcv1["label"], True, ep == 0)
        ref = refout["sgd_v8_trace"][ep]
        assert abs(pr[0] - ref[0]) <= 2e-5 * abs(ref[0])     # float reduction order only
        assert pr[1] == ref[1] and pr[4] == ref[4]
    vals, lens = M.get(keys)
    assert np.array_equal(lens, refout["sgd_v8_final_lens"])
    assert np.array_equal(vals, refout["sgd_v8_final_vals"])   # bit-exact incl. InitV's rand_r stream
    assert (lens > 1).sum() > 100


def test_synthetic_trace_vs_reference(refout):
    kw = parse_kwargs(refout["syn_kwargs"])
    M = O.Oracle(**kw)
    batches = syn_batches(refout)
    t = 0
    for ep in range(3):
        for (o, l, i, v) in batches:
            pr = M.sgd_step(o, i, v, l, True, ep == 0)
            ref = refout["syn_trace"][t]
            assert abs(pr[0] - ref[0]) <= 2e-5 * abs(ref[0])
            assert pr[1] == ref[1]
            t += 1
    vals, lens = M.get(refout["syn_keys"])
    assert np.array_equal(lens, refout["syn_final_lens"])
    assert np.array_equal(vals, refout["syn_final_vals"])


def test_auc_untied(refout):
    assert O.auc(refout["auc_label"], refout["auc_pred"]) == pytest.approx(float(refout["auc_value"]), rel=1e-6)
    assert O.auc(np.ones(5, np.float32), np.arange(5, dtype=np.float32)) == 1.0   # single class -> 1


def test_owner_rule():
    # ps-lite postoffice.cc:127-136: range i = [kMax/S*i, kMax/S*(i+1))
    mx = 0xFFFFFFFFFFFFFFFF
    for S in (1, 2, 3, 4, 8):
        w = mx // S
        for i in range(S):
            assert O.orc().orc_owner(w * i, S) == i
            assert O.orc().orc_owner(w * (i + 1) - 1, S) == i
        assert O.orc().orc_owner(mx, S) == S - 1
        ks = np.array([0, w - 1, w, mx - 1, mx], dtype=np.uint64)
        assert np.array_equal(O.owner(ks, S), [O.orc().orc_owner(int(k), S) for k in ks])
