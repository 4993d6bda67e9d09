#!/usr/bin/env python
"""bench.py -- FM-SGD examples/sec on synthetic Criteo-shaped CSR minibatches (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one minibatch through the whole hot path of SGDLearner::IterateData
(src/sgd/sgd_learner.cc:138-177 of the reference): Pull -> FM forward -> logloss/AUC/penalty ->
FM backward -> Push (FTRL on w, AdaGrad on V).  Workload (default): B=65536 rows x 100 nnz,
feature ids uniform over [0, 1e9), binary values, V_dim=64, V_threshold=0 and l1=0 so that every
key owns a V row (the bandwidth worst case of SURVEY.md 8d, config "S").

Printed JSON line (rank 0):
  value        examples/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the C-ABI call a user makes, every step's inputs copied
               host->device from pinned memory and every step's Progress read back
  roofline     dominant kernel (fused FM forward+backward): algorithmic bytes / event-timed
               duration vs the measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline the reference's own CPU SGD path (oracle/_ref) on a bounded sample, host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--nnz", type=int, default=100)
    ap.add_argument("--vdim", type=int, default=64)
    ap.add_argument("--id-space", type=int, default=10 ** 9)
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "criteo39"])
    ap.add_argument("--working-set", type=int, default=8, help="distinct batches cycled (per rank)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap-auc", action="store_true", help="keep everything on one stream (profiler runs)")
    ap.add_argument("--cpu-rows", type=int, default=4096, help="rows per CPU-baseline sample batch")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------
def reverse_bytes_np(x):
    """nibble reversal of a uint64 array (include/difacto/base.h:39-51 of the reference)"""
    x = x.astype(np.uint64, copy=True)
    x = (x << np.uint64(32)) | (x >> np.uint64(32))
    x = ((x & np.uint64(0x0000FFFF0000FFFF)) << np.uint64(16)) | ((x & np.uint64(0xFFFF0000FFFF0000)) >> np.uint64(16))
    x = ((x & np.uint64(0x00FF00FF00FF00FF)) << np.uint64(8)) | ((x & np.uint64(0xFF00FF00FF00FF00)) >> np.uint64(8))
    x = ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4)) | ((x & np.uint64(0xF0F0F0F0F0F0F0F0)) >> np.uint64(4))
    return x


def gen_raw_batch(args, seed, rows=None):
    """raw (un-localized) CSR<u64> minibatch of the named shape"""
    rng = np.random.default_rng(seed)
    B = rows or args.batch
    if args.workload == "criteo39":
        # 13 "integer" + 26 "categorical" groups, id = (hash << 12) | group (criteo_parser.h:68-88), Zipf-ish hashes
        nnz = 39
        gid = np.tile(np.arange(nnz, dtype=np.uint64), B)
        z = rng.zipf(1.05, B * nnz).astype(np.uint64) % np.uint64(1 << 26)
        ids = (z << np.uint64(12)) | gid
    else:
        nnz = args.nnz
        ids = rng.integers(0, args.id_space, B * nnz).astype(np.uint64)
    off = (np.arange(B + 1, dtype=np.uint64) * np.uint64(nnz))
    lab = np.where(rng.random(B) < 0.25, 1.0, -1.0).astype(np.float32)
    return off, lab, ids


def localize_np(ids):
    """host restatement of Localizer::Compact (src/data/localizer.cc:11-103) with numpy:
    reversed keys, ascending unique, occurrence counts, rank of every nnz"""
    rk = reverse_bytes_np(ids)
    keys, lidx, cnt = np.unique(rk, return_inverse=True, return_counts=True)
    return lidx.astype(np.uint32), keys.astype(np.uint64), cnt.astype(np.float32)


def hyper(args):
    # every key gets a V row: V_threshold=0, l1=0; other values are the defaults
    # of src/sgd/sgd_param.h:94-106
    return dict(V_dim=args.vdim, l1=0.0, l2=0.0, lr=0.01, lr_beta=1.0, V_l2=0.01, V_lr=0.01, V_lr_beta=1.0,
                V_init_scale=0.01, V_threshold=0, seed=0)


# --------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
                for n, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# the reference arm / cpu baseline: the reference's own CPU SGD path (oracle/_ref), bounded sample
# --------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup, nthreads=None, quiet=True):
    from oracle import oracle as O   # test infrastructure: allowed here as the CPU baseline only
    cores = os.cpu_count() or 1
    if nthreads is None:
        nthreads = max(2, min(48, cores))     # loss.h:80-83 caps nthreads at < 50
    kind = "reference" if O.have_ref() else "port"
    kw = hyper(args)
    rows = args.cpu_rows
    nb = 2
    batches = [gen_raw_batch(args, 10_000 + b, rows=rows) for b in range(nb)]
    eng = O.RefOracle(nthreads=nthreads, **kw) if kind == "reference" else O.Oracle(**kw)
    # table warm-up: two passes so that every key of the sample owns a V row (as on the GPU arm)
    for p in range(2):
        for (off, lab, ids) in batches:
            eng.sgd_step(off, ids, None, lab, True, p == 0)
    for t in range(warmup):
        off, lab, ids = batches[t % nb]
        eng.sgd_step(off, ids, None, lab, True, False)
    secs = np.zeros(6, np.float64)
    t0 = time.perf_counter()
    for t in range(steps):
        off, lab, ids = batches[t % nb]
        if kind == "reference":
            eng.sgd_step(off, ids, None, lab, True, False, seconds=secs)
        else:
            eng.sgd_step(off, ids, None, lab, True, False)
    dt = time.perf_counter() - t0
    ex_s = steps * rows / dt
    sample = (f"{steps} steps x {rows} rows x {args.nnz if args.workload == 'synthetic' else 39} nnz "
              f"(same shape/hyper-parameters, table warmed, Localizer::Compact included)")
    return dict(value=ex_s, unit="examples/s", cores=int(nthreads if kind == "reference" else 1), kind=kind,
                sample=sample, ms_per_step=dt / steps * 1e3, host_cores=cores,
                stage_seconds=dict(zip(["localize", "feacnt", "get", "predict", "calcgrad", "update"],
                                       [float(x) for x in secs])))


def main_reference(args, rank, world):
    if rank != 0:
        return
    res = cpu_reference_run(args, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": res["value"], "unit": "examples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, None),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": res["host_cores"], "stage_seconds": res["stage_seconds"],
    }
    print(json.dumps(line))


def metric_name(args):
    return f"FM-SGD examples/sec, synthetic Criteo-shape CSR, V_dim={args.vdim}"


def workload_config(args, extra):
    nnz = args.nnz if args.workload == "synthetic" else 39
    cfg = {"workload": (f"synthetic CSR (BASELINE.json configs[4] / SURVEY 8d 'S'): batch {args.batch} x {nnz} nnz/row, "
                        f"ids uniform over [0,{args.id_space}), binary values, V_dim={args.vdim}, every key owns a V row"
                        if args.workload == "synthetic" else
                        f"criteo-shaped CSR: batch {args.batch} x 39 nnz/row, Zipf ids with 12-bit group id, V_dim={args.vdim}"),
           "global_batch": args.batch * args.gpus, "batch_per_gpu": args.batch, "nnz_per_row": nnz, "V_dim": args.vdim,
           "id_space": args.id_space, "hyper": hyper(args),
           "l2_policy": "inputs_exceed_l2 (each step gathers/updates >1 GB of distinct table rows; "
                        "the working set of batches is cycled, never the same batch twice in a row)"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------
def main_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from difacto_b200 import capi

    if world > 1:
        from difacto_b200 import sharded
        return sharded.bench_main(args, rank, world, local_rank, sys.modules[__name__])

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    kw = hyper(args)
    nb = args.working_set
    nnz = args.nnz if args.workload == "synthetic" else 39

    # ---- synthetic data: raw ids -> localized batches in pinned host memory ----
    t_gen = time.perf_counter()
    host = []
    total_keys = 0
    for b in range(nb):
        off, lab, ids = gen_raw_batch(args, 1 + b)
        lidx, keys, cnt = localize_np(ids)
        total_keys += len(keys)
        hb = dict(off=torch.from_numpy(off).pin_memory(), lab=torch.from_numpy(lab).pin_memory(),
                  lidx=torch.from_numpy(lidx.view(np.int32)).pin_memory(),
                  keys=torch.from_numpy(keys.view(np.int64)).pin_memory(),
                  cnt=torch.from_numpy(cnt).pin_memory(), U=len(keys))
        host.append(hb)
    t_gen = time.perf_counter() - t_gen
    B = args.batch
    N = B * nnz
    U_mean = total_keys / nb

    cap = int(total_keys * 1.05) + 1024
    extra = {}
    if os.environ.get("DFB_L2_FETCH"):
        extra["l2_fetch_granularity"] = int(os.environ["DFB_L2_FETCH"])
    E = capi.Engine(device=local_rank, table_capacity=cap, V_capacity=cap,
                    overlap_auc=0 if args.no_overlap_auc else 1, **extra, **kw)
    ks = E.row_stride()
    # the product's own localizer (Localizer::Compact on the GPU) must agree bit for bit with the harness's numpy one
    _, _, ids0 = gen_raw_batch(args, 1)
    gl, gk, gc = E.localize(host[0]["off"].numpy(), ids0)
    localizer_check = bool(np.array_equal(gl.view(np.int32), host[0]["lidx"].numpy())
                           and np.array_equal(gk.view(np.int64), host[0]["keys"].numpy())
                           and np.array_equal(gc, host[0]["cnt"].numpy()))
    devb = [dict(off=h["off"].to(dev), lab=h["lab"].to(dev), lidx=h["lidx"].to(dev), keys=h["keys"].to(dev),
                 cnt=h["cnt"].to(dev), U=h["U"]) for h in host]
    torch.cuda.synchronize()

    def step_dev(b, with_cnt=False, train=True):
        d = devb[b]
        E.train_step_dev(B, N, d["off"], d["lidx"], None, d["lab"], d["keys"], d["U"], d["cnt"] if with_cnt else None, train)

    # ---- table warm-up (untimed): two passes, after which every key owns a V row ----
    for p in range(2):
        for b in range(nb):
            step_dev(b, with_cnt=(p == 0))
    pr = E.read_progress()
    st = E.table_stats()
    assert st["n_vrows"] == st["n_keys"] <= total_keys, (st, total_keys)   # every key owns a V row

    stream = torch.cuda.ExternalStream(E.stream(), device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)

    # ---- value: device-resident inputs, CUDA events on the engine's stream ----
    for t in range(args.warmup):
        step_dev(t % nb)
    E.sync()
    launches0 = E.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.time()
    ev0.record(stream)
    for t in range(args.steps):
        step_dev((args.warmup + t) % nb)
    ev1.record(stream)
    E.sync()
    torch.cuda.synchronize()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = E.launch_count() - launches0
    prog = E.read_progress()
    value = args.steps * B / (ms * 1e-3)

    # ---- the same steps again with per-stage CUDA events (kernel durations for the roofline) ----
    E.profile(True)
    for t in range(args.steps):
        step_dev((args.warmup + t) % nb)
    E.sync()
    stages = E.profile_read()
    E.profile(False)
    E.read_progress()

    # ---- forward-only launches (validation batches): the pure gather+interaction kernel K1 ----
    E.profile(True)
    for t in range(max(4, args.steps // 2)):
        step_dev(t % nb, train=False)
    E.sync()
    stages_fwd = E.profile_read()
    E.profile(False)
    E.read_progress()

    # ---- e2e: host (pinned) buffers through dfb_train_step_async, Progress read back every step ----
    e2e = None
    if not args.no_e2e:
        def submit(b):
            h = host[b]
            E.train_step_async(B, h["off"], h["lidx"], None, h["lab"], h["keys"], h["U"], None, True)
        for t in range(args.warmup):
            submit(t % nb)
        E.read_progress()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss_sum = 0.0
        for t in range(args.steps):
            submit((args.warmup + t) % nb)
            if t >= 1:
                loss_sum += E.wait_step().loss      # D2H read of step t-1's result while step t runs
        loss_sum += E.wait_step().loss
        E.sync()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = (B + 1) * 8 + N * 4 + B * 4 + int(U_mean) * 8
        e2e = {"value": args.steps * B / dt, "unit": "examples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": 64, "ms_per_step": dt / args.steps * 1e3,
               "api": "dfb_train_step_async + dfb_wait_step (C-ABI), localized CSR + keys from pinned host memory",
               "mean_loss_per_step": loss_sum / args.steps}
    # ---- e2e from RAW ids: Localizer::Compact runs on the GPU inside the timed region ----
    e2e_raw = None
    if not args.no_e2e:
        raw = []
        for b in range(nb):
            off, lab, ids = gen_raw_batch(args, 1 + b)
            raw.append(dict(off=torch.from_numpy(off).pin_memory(), lab=torch.from_numpy(lab).pin_memory(),
                            ids=torch.from_numpy(ids.view(np.int64)).pin_memory()))

        def submit_raw(b, nxt=None):
            r = raw[b]
            E.train_step_raw_async(B, r["off"], r["ids"], None, r["lab"], False, True)
            if nxt is not None:       # start the H2D of the next batch while this step runs
                n_ = raw[nxt]
                E.prefetch_raw(B, n_["off"], n_["ids"], None, n_["lab"])
        for t in range(args.warmup):
            submit_raw(t % nb)
        E.read_progress()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss_sum = 0.0
        for t in range(args.steps):
            submit_raw((args.warmup + t) % nb, (args.warmup + t + 1) % nb if t + 1 < args.steps else None)
            if t >= 1:
                loss_sum += E.wait_step().loss
        loss_sum += E.wait_step().loss
        E.sync()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_raw = {"value": args.steps * B / dt, "unit": "examples/s",
                   "h2d_bytes_per_step": int((B + 1) * 8 + N * 8 + B * 4), "d2h_bytes_per_step": 64 + 16,
                   "ms_per_step": dt / args.steps * 1e3,
                   "api": "dfb_train_step_raw_async (+ dfb_prefetch_raw) + dfb_wait_step: raw uint64 CSR from pinned host memory, "
                          "Localizer::Compact on the GPU, then the fused step"}
    sampler.stop()
    clocks = sampler.summary(wall0, time.time())   # value + stage + forward-only + e2e regions: all under load

    # ---- roofline of the dominant kernel (fused FM forward + backward scatter) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    k = args.vdim
    U = U_mean
    # algorithmic bytes:
    #  K1 gather+interaction: N(4k+8) + 16B
    #  emit variant used in training adds the p*XV rows (4kB + 4B) and the per-nnz row payload (4N)
    #  K2+K3 fused per-key reduce + FTRL/AdaGrad: per key read {entry 16, V|cg 8k, slot/vrow/col 16},
    #     write {entry 16, V|cg 8k}; per nnz read {p*XV row 4k, p 4, payload 4}
    bytes_fwd = N * (4 * k + 8) + 16 * B
    bytes_emit = bytes_fwd + B * (4 * k + 4) + 4 * N
    bytes_upd = U * (2 * (8 * k + 16) + 16) + N * (4 * k + 8)
    upd = stages["update"]
    upd_ms = upd["ms"] / max(upd["count"], 1)
    fm = stages["fm"]
    fm_ms = fm["ms"] / max(fm["count"], 1)
    ach_upd = bytes_upd / (upd_ms * 1e-3) / 1e9 if upd_ms > 0 else 0.0
    ach_fm = bytes_emit / (fm_ms * 1e-3) / 1e9 if fm_ms > 0 else 0.0
    fwd = stages_fwd["fm"]
    fwd_ms = fwd["ms"] / max(fwd["count"], 1)
    ach_fwd = bytes_fwd / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else 0.0
    traffic = None
    try:   # DRAM bytes per launch from the committed ncu --set full capture of the same kernel/config
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")))
        if args.batch == 65536 and nnz == 100 and args.workload == "synthetic":
            e_ = tr.get(f"k_bwd_update<{k}>")
            if e_:
                traffic = e_["dram_read_bytes"] + e_["dram_write_bytes"]
    except Exception:
        pass
    roofline = {"bound": "hbm",
                "kernel": f"k_bwd_update<{k}> (per-key gradient reduce fused with FTRL/AdaGrad; dominant kernel of the step; "
                          "the timed stage also contains the 3 tiny InitV-pass kernels)",
                "achieved": ach_upd, "peak": peak, "unit": "GB/s", "frac": ach_upd / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel_ms": upd_ms, "algorithmic_bytes": int(bytes_upd),
                "gather_interaction": {"kernel": f"k_fm_fast<{k},predict> (gather+interaction, K1 of SURVEY 8d; validation launches)",
                                       "achieved": ach_fwd, "frac": ach_fwd / peak, "kernel_ms": fwd_ms,
                                       "algorithmic_bytes": int(bytes_fwd)},
                "gather_interaction_train": {"kernel": f"k_fm_fast<{k},emit> (K1 + p*XV rows + nnz payload)",
                                             "achieved": ach_fm, "frac": ach_fm / peak, "kernel_ms": fm_ms,
                                             "algorithmic_bytes": int(bytes_emit)}}
    U = U_mean
    bytes_step = (bytes_fwd + (N * (4 * k + 8) + 16 * B + N * 4 * (k + 1)) + U * (4 * (k + 1) + 8 * k + 12)
                  + U * (8 * k + 12))                       # BASELINE.md section 3 full-step model
    step_roof = {"model_bytes": int(bytes_step), "achieved": bytes_step / (ms / args.steps * 1e-3) / 1e9,
                 "frac": bytes_step / (ms / args.steps * 1e-3) / 1e9 / peak}

    cpu = None
    if not args.no_cpu_baseline:
        try:
            res = cpu_reference_run(args, steps=3, warmup=1)
            cpu = {k2: res[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
            cpu["host_cores"] = res["host_cores"]
            res2 = cpu_reference_run(args, steps=2, warmup=1, nthreads=2)
            cpu["value_reference_default_2_threads"] = res2["value"]
        except Exception as e:   # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "examples/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}

    line = {
        "metric": metric_name(args), "value": value, "unit": "examples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, {"unique_keys_per_batch": int(U_mean), "working_set_batches": nb,
                                         "table_keys": int(total_keys), "parallelism": "1 gpu, table resident in HBM"}),
        "roofline": roofline, "step_roofline": step_roof, "cpu_baseline": cpu,
        # headline e2e = from the reader's raw uint64 CSR (Localizer::Compact inside the timed region, as in
        # the reference arm's step); e2e_localized = the same with the batch localized beforehand
        "e2e": e2e_raw if e2e_raw else e2e, "e2e_localized": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
        "stages_ms_per_step": {n: s["ms"] / max(s["count"], 1) for n, s in stages.items()},
        "loss_per_example": prog.loss / max(prog.nrows, 1), "datagen_s": t_gen,
        "gpu_localizer_equals_numpy_localizer": localizer_check,
    }
    print(json.dumps(line))
    E.close()


def main():
    args = parse_args()
    args.warmup = max(args.warmup, 3)      # timing hygiene: never fewer than 3 untimed warm-up steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank, world)
    return main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
