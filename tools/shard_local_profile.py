"""tools/shard_local_profile.py -- the sharded step's kernels under ncu on ONE GPU.

ncu must not wrap a multi-rank command (it replays and serialises kernels), so the per-kernel times of the NVLink-sharded
step are taken from S ranks that live in one process on one device: the same kernels, the same mailbox traffic
(device-local instead of over NVLink), one host thread interleaving the five enqueue phases -- in that order every
device-side wait refers to work that has already run, so serialisation is harmless.  Usage (under gpurun, 1 GPU):

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/shard_launches.csv \
      python tools/shard_local_profile.py --ranks 8 --steps 3
  python tools/shard_local_profile.py --ranks 8 --steps 6 --time     # CUDA-event time per step, no profiler
"""
import argparse
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from difacto_b200 import capi  # noqa: E402
from difacto_b200.sharded import FusedShardedStore  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--nnz", type=int, default=100)
    ap.add_argument("--vdim", type=int, default=64)
    ap.add_argument("--id-space", type=int, default=10 ** 9)
    ap.add_argument("--workload", default="synthetic")
    ap.add_argument("--hyper", default="allV")
    ap.add_argument("--engine-kw", default="")
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    a.gpus = 1
    S, B = a.ranks, a.batch
    N = B * bench.nnz_of(a)
    kw = bench.hyper(a)
    extra = dict(kv.split("=") for kv in a.engine_kw.split(",") if kv)
    dev = torch.device("cuda", 0)
    nb = 2
    host = [bench.gen_raw_set(a, nb, 1 + 1000 * r, torch) for r in range(S)]
    devb = [[dict(off=h["off"].to(dev), lab=h["lab"].to(dev), ids=h["ids"].to(dev)) for h in host[r]] for r in range(S)]
    cap = int(nb * N * 1.15) + 4096       # a shard sees 1/S of the keys of all S workers
    id_bits = int(np.ceil(np.log2(float(max(a.id_space, 2))))) if a.workload == "synthetic" else 64
    engines = [capi.Engine(device=0, table_capacity=cap, V_capacity=cap, id_bits=min(id_bits, 64), shard_timeout_ms=20000,
                           **extra, **kw) for _ in range(S)]
    seg = 0 if a.workload == "synthetic" else N
    FusedShardedStore.connect_local(engines, max_rows=B, max_nnz=N, seg_keys=seg, seg_nnz=seg)

    def step(t, push_cnt):
        for r in range(S):
            d = devb[r][t % nb]
            engines[r].shard_begin_dev(B, N, d["off"], d["ids"], None, d["lab"], push_cnt, True)
        for ph in range(5):
            for r in range(S):
                engines[r].shard_phase(ph)

    for t in range(2 * nb):           # table warm-up (every key reaches its steady state)
        step(t, t < nb)
    for E in engines:
        E.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(a.steps):
        step(t, False)
    for E in engines:
        E.sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if a.time:
        print(f"S={S} local ranks on one GPU: {dt / a.steps * 1e3:.3f} ms per collective step "
              f"({S * B * a.steps / dt / 1e6:.2f} M examples/s on this one GPU)")
    for E in engines:
        E.read_progress()
        E.close()


if __name__ == "__main__":
    main()
