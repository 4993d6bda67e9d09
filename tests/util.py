"""helpers shared by the parity tests (test infrastructure)"""
import numpy as np

from oracle import oracle as O


def rand_batch(rng, B, max_nnz, id_space, valued, min_nnz=0, mult=0x9E3779B97F4A7C15):
    nnzr = rng.integers(min_nnz, max_nnz + 1, B)
    off = np.concatenate([[0], np.cumsum(nnzr)]).astype(np.uint64)
    n = int(off[-1])
    idx = rng.integers(0, id_space, n).astype(np.uint64) * np.uint64(mult)
    val = (rng.random(n).astype(np.float32) * 2 - 0.5) if valued else None
    lab = np.where(rng.random(B) < 0.4, 1.0, -1.0).astype(np.float32)
    return off, lab, idx, val


def localized(batch):
    off, lab, idx, val = batch
    lidx, keys, cnt = O.localize(off, idx)
    return dict(offset=off, label=lab, index=idx, value=val, lidx=lidx, keys=keys, cnt=cnt)


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} of {a.size} differ; max err {err.max():.3e} "
                           f"at {np.argmax(err)} (got {a.ravel()[np.argmax(err)]}, want {b.ravel()[np.argmax(err)]})")


def oracle_state(M, keys):
    """dense view of the oracle's entries for keys: scal[n,4], hasv[n], V[n,k], cg[n,k]"""
    k = M.V_dim
    n = len(keys)
    scal = np.zeros((n, 4), np.float32)
    hasv = np.zeros(n, np.int32)
    V = np.zeros((n, k), np.float32)
    cg = np.zeros((n, k), np.float32)
    for i, key in enumerate(keys):
        e = M.lookup(key)
        if e is None:
            hasv[i] = -1
            continue
        scal[i] = [e["fea_cnt"], e["w"], e["sqrt_g"], e["z"]]
        if e["V"] is not None:
            hasv[i] = 1
            V[i], cg[i] = e["V"], e["cg"]
    return scal, hasv, V, cg
