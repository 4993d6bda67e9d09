"""Sequential simulation of the sharded-store semantics with oracle shards -- TEST INFRASTRUCTURE.

S workers, S owners (shards by the ps-lite range rule).  One step:
  [epoch 0] every owner applies the workers' feature counts as separate Update(kFeaCount) calls, rank order;
  every worker pulls the step-start model of its keys, runs Predict / Evaluate / penalty / AUC / CalcGrad with
  the oracle; every owner applies the workers' gradients as separate Update(kGradient) calls in rank order
  (sgd_updater.cc:74-98).  This is what ShardedStore, PeerShardedStore and the fused dfb_shard_* path must equal.
"""
import numpy as np
import torch

from oracle import oracle as O
from oracle_backend import OracleBackend


def key_owner_np(keys, S):
    keys = np.asarray(keys, dtype=np.uint64)
    width = np.uint64(0xFFFFFFFFFFFFFFFF // S)
    return np.minimum(keys // width, np.uint64(S - 1)).astype(np.int64)


def raw_batch(rank, step, valued, B=96, ids=500, max_nnz=24):
    """a small raw (un-localized) CSR<u64> minibatch; ids spread over the whole reversed key space"""
    rng = np.random.default_rng(100 * rank + step)
    nnzr = rng.integers(0, max_nnz, B)
    off = np.concatenate([[0], np.cumsum(nnzr)]).astype(np.uint64)
    n = int(off[-1])
    idx = rng.integers(0, ids, n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    val = rng.random(n).astype(np.float32) if valued else None
    lab = np.where(rng.random(B) < 0.4, 1.0, -1.0).astype(np.float32)
    return off, idx, val, lab


def localized(raw):
    off, idx, val, lab = raw
    lidx, keys, cnt = O.localize(off, idx)
    return dict(nrows=len(lab), nnz=len(idx), U=len(keys), off=torch.from_numpy(off.view(np.int64).copy()),
                lab=torch.from_numpy(lab), lidx=torch.from_numpy(lidx.view(np.int32).copy()),
                keys=torch.from_numpy(keys.view(np.int64).copy()), cnt=torch.from_numpy(cnt),
                val=torch.from_numpy(val) if val is not None else None)


def simulate(S, steps, kw, batch_fn, cnt_steps=2, train_fn=None):
    """batch_fn(rank, step) -> raw batch.  Returns (shards, workers, per-step per-worker progress snapshots)"""
    shards = [OracleBackend(**kw) for _ in range(S)]
    workers = [OracleBackend(**kw) for _ in range(S)]      # fm_step + progress only
    per_step = []
    for step in range(steps):
        is_train = True if train_fn is None else train_fn(step)
        batches = [localized(batch_fn(r, step)) for r in range(S)]
        owners = [key_owner_np(b["keys"].numpy().view(np.uint64), S) for b in batches]
        if step < cnt_steps:
            for s in range(S):
                for r in range(S):
                    m = owners[r] == s
                    shards[s].feacnt(batches[r]["keys"][m], batches[r]["cnt"][m])
        pulled, grads = [], []
        before = [w.progress.copy() for w in workers]
        for r in range(S):
            b = batches[r]
            U, ks = b["U"], shards[0].ks
            w, hasv, V = torch.zeros(U), torch.zeros(U, dtype=torch.int32), torch.zeros(U, ks)
            for s in range(S):
                m = np.nonzero(owners[r] == s)[0]
                ws, hs, Vs = torch.zeros(len(m)), torch.zeros(len(m), dtype=torch.int32), torch.zeros(len(m), ks)
                shards[s].pull_rows(b["keys"][m], ws, hs, Vs)
                w[m], hasv[m], V[m] = ws, hs, Vs
            pulled.append((w, hasv, V))
            gw, gV = torch.zeros(U), torch.zeros(U, ks)
            workers[r].fm_step(b, w, hasv, V, is_train, gw, gV)
            grads.append((gw, gV))
        per_step.append([workers[r].progress - before[r] for r in range(S)])
        if not is_train:
            continue
        for s in range(S):
            for r in range(S):
                m = np.nonzero(owners[r] == s)[0]
                shards[s].push_rows(batches[r]["keys"][m], grads[r][0][m], pulled[r][1][m], grads[r][1][m])
    return shards, workers, per_step


def all_keys(S, steps, batch_fn):
    ks = []
    for r in range(S):
        for st in range(steps):
            off, idx, _, _ = batch_fn(r, st)
            ks.append(O.localize(off, idx)[1])
    return np.unique(np.concatenate(ks))
