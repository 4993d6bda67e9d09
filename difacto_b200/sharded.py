"""NVLink-sharded model store: difacto's Store::Pull/Push as NCCL all-to-all of the active rows.

The reference shards keys over ps-lite servers by contiguous ranges of the (reversed) key space
(ps-lite/src/postoffice.cc:127-136) and a worker slices its sorted key list per server
(ps-lite/include/ps/kv_app.h:406-460, DefaultSlicer); difacto's own distributed Store is
`LOG(FATAL) << "not implemented"` (src/store/store.cc:9-11).  Here every rank is both a worker (its
own minibatch) and a server (shard `rank` of the table, one difacto_b200 engine), one process per
GPU, `torch.distributed` for the plumbing:

    counts      all_to_all of the per-owner segment sizes                    (S ints)
    keys        all_to_all_v of the sorted key segments                      (8 B / key)
    [feacnt]    all_to_all_v of the counts; owner applies Update(kFeaCount) per source rank
    Pull        owner gathers {w, has_V, V[ks]} rows (dfb_dev_pull_rows) -> all_to_all_v back
    compute     dfb_dev_fm_step on the pulled dense rows: loss/AUC/penalty + complete gradients
    Push        all_to_all_v of {gw, gV[ks]} -> owner applies FTRL/AdaGrad per source rank, in
                rank order (the reference applies every worker's push as a separate Update in
                arrival order, sgd_updater.cc:74-98; rank order makes it deterministic)

Because the batch's keys are sorted, every owner's keys are ONE contiguous segment, so the
concatenation of what comes back from owners 0..S-1 is already in key order: no permutation pass.

The protocol is backend-agnostic: `CudaBackend` drives the sm_100a engine (the product);
tests/ injects a CPU backend over gloo to check the protocol itself against the oracle.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def key_owner_np(keys, S):
    """vectorised dfb_key_owner: min(S-1, key // (UINT64_MAX // S)) on reversed keys"""
    keys = np.asarray(keys, dtype=np.uint64)
    width = np.uint64(0xFFFFFFFFFFFFFFFF // S)
    return np.minimum(keys // width, np.uint64(S - 1)).astype(np.int64)


def shard_bounds_np(sorted_keys, S):
    """segment boundaries [S+1] of a sorted key array (DefaultSlicer's lower_bound per range)"""
    keys = np.asarray(sorted_keys, dtype=np.uint64)
    width = 0xFFFFFFFFFFFFFFFF // S
    cuts = np.array([width * i for i in range(1, S)], dtype=np.uint64)
    inner = np.searchsorted(keys, cuts, side="left")
    return np.concatenate([[0], inner, [len(keys)]]).astype(np.int64)


class CudaBackend:
    """one shard = one difacto_b200 engine; all arrays are torch CUDA tensors"""

    def __init__(self, engine, device):
        self.E = engine
        self.device = device
        self.ks = engine.row_stride()
        self.V_dim = engine.V_dim
        self.stream = torch.cuda.ExternalStream(engine.stream(), device=device)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def feacnt(self, keys, cnt):
        if len(keys):
            self.E.dev_feacnt(keys, len(keys), cnt)

    def pull_rows(self, keys, w, hasv, V):
        if len(keys):
            self.E.dev_pull_rows(keys, len(keys), w, hasv, V)

    def fm_step(self, batch, w, hasv, V, is_train, gw, gV):
        self.E.dev_fm_step(batch["nrows"], batch["nnz"], batch["off"], batch["lidx"], batch.get("val"),
                           batch["lab"], batch["U"], w, hasv, V, is_train, gw, gV)

    def push_rows(self, keys, gw, hasv, gV):
        if len(keys):
            self.E.dev_push_rows(keys, len(keys), gw, hasv, gV)

    def read_progress(self):
        return self.E.read_progress()


class ShardedStore:
    """Store + the worker's minibatch step over S ranks"""

    def __init__(self, backend, group=None):
        self.b = backend
        self.group = group
        self.S = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ks = backend.ks
        self._buf = {}
        self.timers = None      # set to {} to collect per-phase CUDA-event timings (device tensors only)
        self._marks = []

    def _mark(self, name):
        if self.timers is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._marks.append((name, ev))

    def flush_timers(self):
        """fold the recorded marks into self.timers[name] = [total_ms, count] (synchronises)"""
        if self.timers is None or not self._marks:
            return self.timers
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(self._marks[:-1], self._marks[1:]):
            if n1 == "begin":
                continue
            t = self.timers.setdefault(n1, [0.0, 0])
            t[0] += e0.elapsed_time(e1)
            t[1] += 1
        self._marks = []
        return self.timers

    def _get(self, name, n, shape_tail, dtype):
        t = self._buf.get(name)
        need = (n,) + shape_tail
        if t is None or t.shape[0] < n or t.dtype != dtype:
            cap = max(int(n * 1.2) + 16, 16)
            t = self.b.empty((cap,) + shape_tail, dtype)
            self._buf[name] = t
        return t[:n]

    def _a2a(self, out, inp, out_splits, in_splits):
        if self.S == 1:
            out.copy_(inp)
        else:
            dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)

    def step(self, batch, is_train=True, push_cnt=False):
        """one minibatch of SGDLearner::IterateData on this rank's batch; collective over ranks.

        batch: dict(nrows, nnz, U, off, lidx, val|None, lab, keys, cnt|None, bounds[S+1])"""
        S, U, ks = self.S, batch["U"], self.ks
        self._mark("begin")
        bounds = batch["bounds"]
        send = [int(bounds[i + 1] - bounds[i]) for i in range(S)]
        if S == 1:
            recv = list(send)
        else:
            t_send = torch.tensor(send, dtype=torch.int64, device=batch["keys"].device)
            t_recv = torch.empty_like(t_send)
            dist.all_to_all_single(t_recv, t_send, group=self.group)
            recv = [int(x) for x in t_recv.tolist()]
        R = sum(recv)
        seg = np.concatenate([[0], np.cumsum(recv)]).astype(np.int64)
        keys_r = self._get("keys_r", R, (), torch.int64)
        self._a2a(keys_r, batch["keys"], recv, send)
        self._mark("a2a_counts_keys")
        if push_cnt:
            cnt_r = self._get("cnt_r", R, (), torch.float32)
            self._a2a(cnt_r, batch["cnt"], recv, send)
            for src in range(S):    # one Update(kFeaCount) per worker, rank order
                a, b = int(seg[src]), int(seg[src + 1])
                self.b.feacnt(keys_r[a:b], cnt_r[a:b])
        # ---- Pull ----
        w_r = self._get("w_r", R, (), torch.float32)
        hasv_r = self._get("hasv_r", R, (), torch.int32)
        V_r = self._get("V_r", R, (ks,), torch.float32)
        self.b.pull_rows(keys_r, w_r, hasv_r, V_r)
        self._mark("owner_gather")
        w = self._get("w", U, (), torch.float32)
        hasv = self._get("hasv", U, (), torch.int32)
        V = self._get("V", U, (ks,), torch.float32)
        self._a2a(w, w_r, send, recv)
        self._a2a(hasv, hasv_r, send, recv)
        if ks:
            self._a2a(V, V_r, send, recv)
        self._mark("a2a_pull")
        # ---- Predict / Evaluate / CalcGrad ----
        gw = self._get("gw", U, (), torch.float32)
        gV = self._get("gV", U, (ks,), torch.float32)
        self.b.fm_step(batch, w, hasv, V, is_train, gw, gV)
        self._mark("worker_fm")
        if not is_train:
            return
        # ---- Push(kGradient) ----
        gw_r = self._get("gw_r", R, (), torch.float32)
        gV_r = self._get("gV_r", R, (ks,), torch.float32)
        self._a2a(gw_r, gw, recv, send)
        if ks:
            self._a2a(gV_r, gV, recv, send)
        self._mark("a2a_push")
        for src in range(S):        # one Update(kGradient) per worker, rank order
            a, b = int(seg[src]), int(seg[src + 1])
            self.b.push_rows(keys_r[a:b], gw_r[a:b], hasv_r[a:b], gV_r[a:b])
        self._mark("owner_update")


class PeerShardedStore(ShardedStore):
    """Same protocol, but the two row exchanges are fused into the kernels that produce the rows:
    the owner's gather kernel stores {w, has_V, V} straight into the requester's pull buffer and the
    worker's gradient kernel stores {gw, gV} straight into the owner's receive buffer, through
    peer pointers over NVLink (CUDA IPC, one process per GPU).  NCCL carries only the key lists,
    the count matrix and two tiny barriers per step.  Needs the engine's sorted scatter path
    (V_dim in {8,16,32,64,128}); otherwise use ShardedStore."""

    def __init__(self, backend, max_keys, max_recv_keys=None, group=None):
        super().__init__(backend, group)
        E = backend.E
        S, ks = self.S, self.ks
        self.Umax = int(max_keys)
        self.Rmax = int(max_recv_keys if max_recv_keys is not None else max_keys * 1.5 + 1024)
        al = lambda n: (int(n) + 255) // 256 * 256
        # pull buffer (owners store into it): w | hasv | V
        self.off_hasv = al(self.Umax * 4)
        self.off_V = self.off_hasv + al(self.Umax * 4)
        pull_bytes = self.off_V + al(self.Umax * ks * 4)
        # push buffer (workers store into it): gw | gV
        self.off_gV = al(self.Rmax * 4)
        push_bytes = self.off_gV + al(self.Rmax * ks * 4)
        self.pull_ptr, pull_h = E.peer_alloc(pull_bytes)
        self.push_ptr, push_h = E.peer_alloc(push_bytes)
        handles = [None] * S
        if S > 1:
            dist.all_gather_object(handles, (pull_h, push_h), group=self.group)
        self.pull_peer, self.push_peer = [], []
        for r in range(S):
            if r == self.rank:
                self.pull_peer.append(self.pull_ptr)
                self.push_peer.append(self.push_ptr)
            else:
                self.pull_peer.append(E.peer_open(handles[r][0]))
                self.push_peer.append(E.peer_open(handles[r][1]))
        dev = backend.device
        self._tick = torch.zeros(1, device=dev)
        self.hasv_r = torch.empty(self.Rmax, dtype=torch.int32, device=dev)
        if S > 1:
            dist.barrier(group=self.group)

    def _barrier(self):
        if self.S > 1:
            dist.all_reduce(self._tick, group=self.group)

    def step(self, batch, is_train=True, push_cnt=False):
        S, U, ks, me, E = self.S, batch["U"], self.ks, self.rank, self.b.E
        self._mark("begin")
        bounds = [int(x) for x in batch["bounds"]]
        send = [bounds[i + 1] - bounds[i] for i in range(S)]
        dev = batch["keys"].device
        if S == 1:
            M = [send]
        else:
            t_send = torch.tensor(send, dtype=torch.int64, device=dev)
            t_all = torch.empty(S * S, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(t_all, t_send, group=self.group)
            M = t_all.view(S, S).tolist()          # M[r][s] = keys rank r sends to owner s
        recv = [int(M[r][me]) for r in range(S)]
        R = sum(recv)
        # every rank sees the whole count matrix, so a buffer overflow anywhere is detected by all ranks
        # in the same step (a local assert on one rank would leave the others waiting in a collective)
        worst_R = max(sum(int(M[r][s]) for r in range(S)) for s in range(S))
        worst_U = max(sum(int(x) for x in M[r]) for r in range(S))
        if worst_R > self.Rmax or worst_U > self.Umax:
            raise RuntimeError(f"peer buffers too small: a rank receives {worst_R} keys (max_recv_keys={self.Rmax}) "
                               f"or sends {worst_U} keys (max_keys={self.Umax}); recreate PeerShardedStore with larger sizes")
        seg = [0]
        for r in range(S):
            seg.append(seg[-1] + recv[r])
        keys_r = self._get("keys_r", R, (), torch.int64)
        self._a2a(keys_r, batch["keys"], recv, send)
        self._mark("a2a_counts_keys")
        if push_cnt:
            cnt_r = self._get("cnt_r", R, (), torch.float32)
            self._a2a(cnt_r, batch["cnt"], recv, send)
            for src in range(S):
                self.b.feacnt(keys_r[seg[src]:seg[src + 1]], cnt_r[seg[src]:seg[src + 1]])
        # ---- Pull: gather + store into the requester's buffer (its key order = owner order) ----
        for ph in range(S):       # ring order: in phase ph every rank stores into a different peer
            src = (me + ph) % S
            n = recv[src]
            if n == 0:
                continue
            off = sum(int(M[src][s]) for s in range(me))     # = requester's bounds[me]
            base = self.pull_peer[src]
            E.dev_pull_rows_peer(keys_r[seg[src]:seg[src + 1]], n, base + off * 4, base + self.off_hasv + off * 4,
                                 base + self.off_V + off * ks * 4, self.hasv_r[seg[src]:seg[src + 1]])
        self._barrier()
        self._mark("pull_gather_store")
        w, hasv, V = self.pull_ptr, self.pull_ptr + self.off_hasv, self.pull_ptr + self.off_V
        # ---- Predict / Evaluate / CalcGrad; gradient rows stored into the owners' buffers ----
        if not is_train:
            gw = self._get("gw", U, (), torch.float32)
            gV = self._get("gV", U, (ks,), torch.float32)
            E.dev_fm_step(batch["nrows"], batch["nnz"], batch["off"], batch["lidx"], batch.get("val"), batch["lab"], U,
                          w, hasv, V, False, gw, gV)
            self._barrier()     # nobody overwrites this pull buffer before every worker is done with it
            self._mark("worker_fm")
            return
        pgw, pgV = [], []
        for s in range(S):
            off = sum(int(M[r][s]) for r in range(me))        # where my segment starts at owner s
            pgw.append(self.push_peer[s] + off * 4)
            pgV.append(self.push_peer[s] + self.off_gV + off * ks * 4)
        E.dev_fm_step_peer(batch["nrows"], batch["nnz"], batch["off"], batch["lidx"], batch.get("val"), batch["lab"], U,
                           w, hasv, V, bounds, pgw, pgV, first_seg=me)
        self._barrier()
        self._mark("worker_fm_store")
        # ---- Push(kGradient): one Update per worker, rank order ----
        for src in range(S):
            n = recv[src]
            if n == 0:
                continue
            a = seg[src]
            E.dev_push_rows(keys_r[a:a + n], n, self.push_ptr + a * 4, self.hasv_r[a:a + n],
                            self.push_ptr + self.off_gV + a * ks * 4)
        self._mark("owner_update")


class FusedShardedStore:
    """The NVLink-sharded store behind the C-ABI (dfb_shard_*, csrc/shard.cu + kernels_shard.cu): the product's
    multi-GPU path.  Unlike ShardedStore / PeerShardedStore above (which move the k-wide rows of the active keys
    to the workers and the gradient rows back, Store::Pull / Push as the reference's ps-lite does), the rows stay
    on their owner: owners compute the partial FM interaction sums of every worker's rows and ship (k+2) floats
    per EXAMPLE; the whole protocol (GPU localizer, slicing, peer stores, step-counter flags) runs on the
    device.  This class only does the plumbing a host has to do once: exchange the mailbox handles (CUDA IPC,
    one process per GPU) through torch.distributed, or wire engines of the same process together."""

    def __init__(self, engine, max_rows, max_nnz, group=None, seg_keys=0, seg_nnz=0):
        self.E = engine
        self.group = group
        self.S = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.mailbox_bytes = engine.shard_init(self.rank, self.S, int(max_rows), int(max_nnz), int(seg_keys), int(seg_nnz))
        ptr, handle = engine.shard_export()
        if self.S > 1:
            handles = [None] * self.S
            dist.all_gather_object(handles, handle, group=group)
            peers = [ptr if r == self.rank else engine.peer_open(handles[r]) for r in range(self.S)]
            engine.shard_connect(peers)
            dist.barrier(group=group)      # nobody starts storing into a mailbox that is not mapped everywhere yet

    @staticmethod
    def connect_local(engines, max_rows, max_nnz, seg_keys=0, seg_nnz=0):
        """N engines of ONE process (one per GPU, or several on one GPU in tests): rank = position in the list"""
        S = len(engines)
        for r, E in enumerate(engines):
            E.shard_init(r, S, int(max_rows), int(max_nnz), int(seg_keys), int(seg_nnz))
        ptrs = [E.shard_export()[0] for E in engines]
        for E in engines:
            E.shard_connect(ptrs)

    def step_dev(self, nrows, nnz, d_off, d_ids, d_val, d_lab, is_train=True, push_cnt=False):
        """collective: one minibatch of raw CSR<uint64> per rank, arrays already on this rank's GPU"""
        self.E.shard_step_dev(nrows, nnz, d_off, d_ids, d_val, d_lab, push_cnt, is_train)

    def step_host(self, nrows, off, ids, val, lab, is_train=True, push_cnt=False):
        """collective: the same from (pinned) host arrays; H2D is double-buffered against the previous step"""
        self.E.shard_step_async(nrows, off, ids, val, lab, push_cnt, is_train)


# ---------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): one process per GPU, launched by torch.distributed.run
# ---------------------------------------------------------------------------------------------
def bench_main(args, rank, world, local_rank, benchmod):
    import json
    from difacto_b200 import capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    kw = benchmod.hyper(args)
    nb = args.working_set
    nnz = args.nnz if args.workload == "synthetic" else 39
    B = args.batch
    N = B * nnz

    # per-rank synthetic batches (data parallel: every rank has its own file part, sgd_learner.cc:78-89)
    host = []
    for b in range(nb):
        off, lab, ids = benchmod.gen_raw_batch(args, 1 + b + 1000 * rank)
        lidx, keys, cnt = benchmod.localize_np(ids)
        host.append(dict(nrows=B, nnz=N, U=len(keys), bounds=shard_bounds_np(keys, world),
                         off=torch.from_numpy(off.view(np.int64)).pin_memory(),
                         lab=torch.from_numpy(lab).pin_memory(),
                         lidx=torch.from_numpy(lidx.view(np.int32)).pin_memory(),
                         keys=torch.from_numpy(keys.view(np.int64)).pin_memory(),
                         cnt=torch.from_numpy(cnt).pin_memory()))
    U_mean = float(np.mean([h["U"] for h in host]))
    # this shard sees about (all ranks' keys) / world distinct keys
    cap = int(nb * U_mean * 1.15) + 4096
    E = capi.Engine(device=local_rank, table_capacity=cap, V_capacity=cap, **kw)
    backend = CudaBackend(E, dev)
    use_p2p = os.environ.get("DFB_SHARDED", "p2p") == "p2p" and E.V_dim in (8, 16, 32, 64, 128)
    store = None
    if use_p2p:
        Umax = int(max(h["U"] for h in host) * 1.02) + 1024
        ok = 1
        try:
            store = PeerShardedStore(backend, max_keys=Umax, max_recv_keys=int(Umax * 1.3))
        except Exception as e:      # e.g. CUDA IPC unavailable: every rank falls back together
            print(f"[rank {rank}] peer store unavailable ({e!r}); using NCCL all_to_all", flush=True)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use_p2p = bool(flag.item())
    if not use_p2p:
        store = ShardedStore(backend)

    def to_dev(h):
        d = dict(h)
        for k in ("off", "lab", "lidx", "keys", "cnt"):
            d[k] = h[k].to(dev, non_blocking=True)
        return d

    devb = [to_dev(h) for h in host]
    torch.cuda.synchronize()

    with torch.cuda.stream(backend.stream):
        for p in range(2):      # table warm-up: afterwards every key owns a V row
            for b in range(nb):
                store.step(devb[b], True, push_cnt=(p == 0))
        E.read_progress()
        for t in range(args.warmup):
            store.step(devb[t % nb], True)
        E.sync()
        dist.barrier()
        torch.cuda.synchronize()
        launches0 = E.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = benchmod.ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        wall0 = time.time()
        ev0.record(backend.stream)
        for t in range(args.steps):
            store.step(devb[(args.warmup + t) % nb], True)
        ev1.record(backend.stream)
        E.sync()
        torch.cuda.synchronize()
        dist.barrier()
        wall1 = time.time()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        launches = E.launch_count() - launches0
        prog = E.read_progress()

        # ---- per-phase CUDA-event timings (separate region) ----
        store.timers = {}
        for t in range(max(4, args.steps // 2)):
            store.step(devb[(args.warmup + t) % nb], True)
        phases = {k2: v2[0] / max(v2[1], 1) for k2, v2 in store.flush_timers().items()}
        store.timers = None
        E.read_progress()

        # ---- e2e: per-step H2D of the localized batch from pinned memory + Progress read back ----
        e2e = None
        if not args.no_e2e:
            # double-buffered input: the H2D copy of batch t+1 runs on a copy stream while step t computes
            copy_stream = torch.cuda.Stream(device=dev)
            main = backend.stream

            def prefetch(h):
                with torch.cuda.stream(copy_stream):
                    d = to_dev(h)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                return d, ev

            def run_e2e(nsteps, first):
                loss = 0.0
                nxt = prefetch(host[first % nb])
                for t in range(nsteps):
                    cur, ev = nxt
                    if t + 1 < nsteps:
                        nxt = prefetch(host[(first + t + 1) % nb])
                    main.wait_event(ev)
                    store.step(cur, True)
                    for k2 in ("off", "lab", "lidx", "keys", "cnt"):
                        cur[k2].record_stream(main)
                    loss += E.read_progress().loss      # D2H of the step's result
                return loss

            run_e2e(args.warmup, 0)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loss_sum = run_e2e(args.steps, args.warmup)
            torch.cuda.synchronize()
            dist.barrier()
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.item())
            h2d = (B + 1) * 8 + N * 4 + B * 4 + int(U_mean) * 8
            e2e = {"value": args.steps * B * world / dt, "unit": "examples/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": 64, "ms_per_step": dt / args.steps * 1e3,
                   "api": "ShardedStore.step (torch.distributed all_to_all + C-ABI dfb_dev_*), localized CSR + keys "
                          "from pinned host memory every step"}
    if rank == 0:
        sampler.stop()
        k = args.vdim
        a2a_bytes = 2 * U_mean * (world - 1) / world * (8 + 4 * (k + 1))     # BASELINE.md section 3, per direction
        nvl = 770.0   # measured peer copy GB/s per direction (B200_PROFILING.md)
        line = {
            "metric": benchmod.metric_name(args), "value": args.steps * B * world / (ms * 1e-3), "unit": "examples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": benchmod.workload_config(args, {
                "unique_keys_per_batch": int(U_mean), "working_set_batches": nb,
                "parallelism": f"dp{world} minibatches x table sharded by reversed-key range over {world} GPUs "
                               "(ps-lite rule), Pull/Push of the active rows = "
                               + ("peer stores over NVLink fused into the gather / gradient kernels (CUDA IPC)"
                                  if use_p2p else "NCCL all_to_all")}),
            "roofline": {"bound": "nvlink", "kernel": ("k_gather_rows -> peer pull buffer + k_bwd_update<dense> -> peer push buffer "
                                                       "(fused compute + NVLink stores)" if use_p2p else
                                                       "all_to_all of active rows (pull + push)"),
                         "achieved": a2a_bytes / (ms / args.steps * 1e-3) / 1e9, "peak": nvl, "unit": "GB/s",
                         "frac": a2a_bytes / (ms / args.steps * 1e-3) / 1e9 / nvl, "traffic": None,
                         "algorithmic_bytes": int(a2a_bytes),
                         "note": "bytes that must cross NVLink per GPU per direction per step / whole step time"},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(wall0, wall1), "loss_per_example": prog.loss / max(prog.nrows, 1),
            "phases_ms_per_step_rank0": phases,
        }
        print(json.dumps(line), flush=True)
    # Tensors (device and pinned host) that were used on the engine's stream record events on it
    # when they are freed, which at interpreter teardown can happen after the stream is gone:
    # finish cleanly and leave without running destructors.
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    import sys
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)
