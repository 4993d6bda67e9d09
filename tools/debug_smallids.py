"""debug: the sharded store with SMALL feature ids (the localizer learns a narrow bit range, 32-bit sort keys) vs the oracle simulation"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from difacto_b200 import capi
from difacto_b200.sharded import FusedShardedStore
from shard_sim import all_keys, key_owner_np, simulate
from util import oracle_state

KW = dict(V_dim=16, l1=0.01, l2=0.01, lr=0.1, V_lr=0.05, V_threshold=2, V_l2=0.01, V_init_scale=0.01, seed=0)
S, STEPS = 2, 5

def batch_fn(rank, step):
    rng = np.random.default_rng(1000 * rank + step)
    B = 200
    rows = [np.unique(rng.integers(1, 400, rng.integers(5, 25))) for _ in range(B)]
    off = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint64)
    idx = (np.concatenate(rows).astype(np.uint64) * np.uint64(7919))
    val = rng.random(len(idx)).astype(np.float32)
    lab = np.where(rng.random(B) < 0.5, 1.0, -1.0).astype(np.float32)
    return off, idx, val, lab

TWO = os.environ.get("DEVS") == "2"
engines = [capi.Engine(device=(r if TWO else 0), table_capacity=1 << 16, shard_timeout_ms=8000, **KW) for r in range(S)]
FusedShardedStore.connect_local(engines, max_rows=65536, max_nnz=200000)
for step in range(STEPS):
    for r in range(S):
        off, idx, val, lab = batch_fn(r, step)
        engines[r].shard_begin_async(len(lab), off, idx, val, lab, push_cnt=True, is_train=True)
    for ph in range(5):
        for r in range(S):
            engines[r].shard_phase(ph)
    pr = [engines[r].wait_step() for r in range(S)]
    print("step", step, [round(p.loss, 4) for p in pr], flush=True)
shards, workers, per_step = simulate(S, STEPS, KW, batch_fn, cnt_steps=STEPS)
print("oracle", [[round(float(x[0]), 4) for x in st] for st in per_step])
keys = all_keys(S, STEPS, batch_fn); own = key_owner_np(keys, S)
for s in range(S):
    scal, hasv, V, cg = engines[s].read_entries(keys[own == s])
    oscal, ohasv, oV, ocg = oracle_state(shards[s].M, keys[own == s])
    print("shard", s, "keys", (own == s).sum(), "hasv equal", np.array_equal(hasv, ohasv), "max|dw|", np.abs(scal[:, 1] - oscal[:, 1]).max(), "max|dV|", np.abs(V - oV).max())
