"""world_size-2 gloo tests of the sharded store protocol (CPU; the compute backend is the oracle).

Checks that Pull/Push through all_to_all_v over 2 ranks equals a sequential simulation of the same
parameter-server semantics: every worker pulls the step-start model, owners apply the workers'
pushes as separate Updates in rank order (sgd_updater.cc:74-98), keys live on the shard given by
the ps-lite range rule (postoffice.cc:127-136)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from difacto_b200.sharded import ShardedStore, key_owner_np, shard_bounds_np  # noqa: E402

KW = dict(V_dim=4, l1=0.05, l2=0.01, lr=0.2, V_lr=0.1, V_threshold=1, V_l2=0.01, V_init_scale=0.2, seed=3)
if os.environ.get("DFB_TEST_VDIM") is not None:      # the spawned workers inherit the variant through the environment
    KW["V_dim"] = int(os.environ["DFB_TEST_VDIM"])
S = 2
STEPS = 6


def make_batch(rank, step, valued):
    rng = np.random.default_rng(100 * rank + step)
    B = 40
    nnzr = rng.integers(0, 12, B)
    off = np.concatenate([[0], np.cumsum(nnzr)]).astype(np.uint64)
    n = int(off[-1])
    # ids chosen so that the reversed keys spread over both halves of the key space
    idx = rng.integers(0, 150, n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    val = rng.random(n).astype(np.float32) if valued else None
    lab = np.where(rng.random(B) < 0.4, 1.0, -1.0).astype(np.float32)
    lidx, keys, cnt = O.localize(off, idx)
    d = dict(nrows=B, nnz=n, U=len(keys), bounds=shard_bounds_np(keys, S),
             off=torch.from_numpy(off.view(np.int64).copy()), lab=torch.from_numpy(lab),
             lidx=torch.from_numpy(lidx.view(np.int32).copy()), keys=torch.from_numpy(keys.view(np.int64).copy()),
             cnt=torch.from_numpy(cnt), val=torch.from_numpy(val) if valued else None)
    return d


def simulate():
    """sequential reference of the protocol with S oracle shards"""
    from oracle_backend import OracleBackend
    shards = [OracleBackend(**KW) for _ in range(S)]
    workers = [OracleBackend(**KW) for _ in range(S)]      # only used for fm_step + progress
    for step in range(STEPS):
        batches = [make_batch(r, step, step % 2 == 0) for r in range(S)]
        owners = [key_owner_np(b["keys"].numpy().view(np.uint64), S) for b in batches]
        if step < 2:
            for s in range(S):
                for r in range(S):
                    m = owners[r] == s
                    shards[s].feacnt(batches[r]["keys"][m], batches[r]["cnt"][m])
        pulled = []
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
        grads = []
        for r in range(S):
            b = batches[r]
            w, hasv, V = pulled[r]
            gw, gV = torch.zeros(b["U"]), torch.zeros(b["U"], shards[0].ks)
            workers[r].fm_step(b, w, hasv, V, True, gw, gV)
            grads.append((gw, gV))
        for s in range(S):
            for r in range(S):
                m = np.nonzero(owners[r] == s)[0]
                shards[s].push_rows(batches[r]["keys"][m], grads[r][0][m], pulled[r][1][m], grads[r][1][m])
    return shards, workers


def worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=S)
    from oracle_backend import OracleBackend
    be = OracleBackend(**KW)
    store = ShardedStore(be)
    for step in range(STEPS):
        b = make_batch(rank, step, step % 2 == 0)
        store.step(b, True, push_cnt=step < 2)
    # dump this shard's model: all keys it may own
    allkeys = np.unique(np.concatenate([make_batch(r, st, False)["keys"].numpy().view(np.uint64)
                                        for r in range(S) for st in range(STEPS)]))
    vals, lens = be.M.get(allkeys[key_owner_np(allkeys, S) == rank])
    np.savez(out.format(rank=rank), vals=vals, lens=lens, progress=be.progress, size=be.M.size())
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_rule_matches_ps_lite():
    rng = np.random.default_rng(0)
    keys = np.sort(rng.integers(0, 2 ** 63, 5000).astype(np.uint64) * np.uint64(2) + np.uint64(1))
    for nshard in (1, 2, 3, 4, 8):
        own = key_owner_np(keys, nshard)
        assert np.array_equal(own, [O.orc().orc_owner(int(k), nshard) for k in keys])
        bounds = shard_bounds_np(keys, nshard)
        assert bounds[0] == 0 and bounds[-1] == len(keys)
        for s in range(nshard):
            assert np.all(own[bounds[s]:bounds[s + 1]] == s)     # contiguous segments in owner order
        try:
            from difacto_b200 import capi
            assert np.array_equal(capi.shard_bounds(keys, nshard), bounds)     # the C-ABI does the same
            assert [capi.key_owner(int(k), nshard) for k in keys[:200]] == list(own[:200])
        except ImportError:
            pass


@pytest.mark.parametrize("V_dim", [4, 0])
def test_two_rank_protocol_equals_sequential_simulation(tmp_path, V_dim, monkeypatch):
    monkeypatch.setenv("DFB_TEST_VDIM", str(V_dim))
    KW["V_dim"] = V_dim
    out = str(tmp_path / "shard{rank}.npz")
    mp.spawn(worker, args=(free_port(), out), nprocs=S, join=True)
    shards, workers = simulate()
    allkeys = np.unique(np.concatenate([make_batch(r, st, False)["keys"].numpy().view(np.uint64)
                                        for r in range(S) for st in range(STEPS)]))
    own = key_owner_np(allkeys, S)
    total_v = 0
    for s in range(S):
        got = np.load(out.format(rank=s))
        vals, lens = shards[s].M.get(allkeys[own == s])
        assert np.array_equal(got["lens"], lens)
        assert np.array_equal(got["vals"], vals)            # bit-identical model on every shard
        assert np.array_equal(got["progress"], workers[s].progress)
        total_v += int((lens > 1).sum())
        # nothing leaked to the wrong shard: the other shard's keys are absent
        for key in allkeys[own != s][:50]:
            assert shards[s].M.lookup(key) is None
    assert total_v > 10 or V_dim == 0
    assert min((own == s).sum() for s in range(S)) > 10     # both shards actually own keys
