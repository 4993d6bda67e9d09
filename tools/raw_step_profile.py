"""a few raw-id fused steps (GPU localizer inside) for ncu launch lists: python tools/raw_step_profile.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difacto_b200 import capi
B, NNZ, K = 65536, 100, 64
rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)
E = capi.Engine(V_dim=K, l1=0, l2=0, V_threshold=0, table_capacity=1 << 25, overlap_auc=0)
bs = []
for b in range(3):
    ids = torch.from_numpy(rng.integers(0, 10 ** 9, B * NNZ).astype(np.int64)).to(dev)
    off = torch.from_numpy((np.arange(B + 1) * NNZ).astype(np.int64)).to(dev)
    lab = torch.from_numpy(np.where(rng.random(B) < 0.25, 1.0, -1.0).astype(np.float32)).to(dev)
    bs.append((off, ids, lab))
for p in range(2):
    for (off, ids, lab) in bs:
        E.train_step_raw_dev(B, B * NNZ, off, ids, None, lab, p == 0, True)
E.sync()
for t in range(6):
    off, ids, lab = bs[t % 3]
    E.train_step_raw_dev(B, B * NNZ, off, ids, None, lab, False, True)
E.sync()
print(E.read_progress().as_dict())
