"""Condense the round-2 gpurun_out/ captures (ncu launch list, ncu --set full reports, A/B logs, bench lines, NVLink
counters) into the tracked summaries under profiles/.  Run here (no GPU needed): python tools/make_profiles_r2.py"""
import collections
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
UNIT_GB = {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}


def bench_line(fn):
    if not os.path.exists(os.path.join(G, fn)):
        return None
    for l in open(os.path.join(G, fn)):
        if l.startswith("{"):
            return json.loads(l)
    return None


def nice(name):
    return re.sub(r'\(.*', '', name).replace('void ', '').replace('dfb::<unnamed>::', '').replace('unnamed>::', '')[:70]


# ------------------------------------------------------------------ launch list
def launches():
    lines = [l for l in open(os.path.join(G, "launches_r2.csv")) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except Exception:
            continue
        unit = row['Metric Unit']
        ms = v / 1e6 if unit.startswith('n') else v / 1e3 if unit.startswith('u') else v
        a = agg.setdefault(nice(row['Kernel Name']), [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    out = ["# ncu launch list, round 2 (B200, sm_100a)", "",
           "Command (under gpurun, 1 GPU): `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py "
           "--steps 4 --warmup 3 --working-set 4 --no-e2e --no-cpu-baseline --no-sweep --no-overlap-auc` (tools/profile_r2.sh; raw csv: "
           "gpurun_out/launches_r2.csv, not tracked).", "",
           "The window covers engine creation, the table warm-up (two passes over 4 batches: the InitV / feature-count kernels do real "
           "work only there), 3 warm-up + 4 timed steps and the first profiled steps.  Times under ncu are cold-cache and serialised: "
           "compare SHARES, not absolutes.  Workload: B=65536 x 100 nnz raw uint64 ids U[0,1e9), V_dim=64, every key owns a V row.", "",
           "| kernel | launches | mean ms | share of window |", "|---|---:|---:|---:|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append(f"| `{n}` | {c} | {t / c:.4f} | {100 * t / tot:.1f}% |")
    d = bench_line("bench_full_final.json") or bench_line("bench_full1.json")
    if d:
        st = d["stages_ms_per_step"]
        tot2 = sum(v for k, v in st.items() if k != "auc")
        out += ["", "## The same step timed live with CUDA events inside bench.py (not under the profiler)", "",
                "`python bench.py` profiled region (kernels alone on one stream), ms/step: "
                + ", ".join(f"{k} {v:.3f}" for k, v in st.items())
                + f"; unprofiled step {d['ms_per_step']:.3f} ms = {d['value'] / 1e6:.1f} M examples/s (localizer and AUC overlapped on side streams).", "",
                f"* dominant kernel `k_bwd_update<64,0,1,0>`: {100 * st['update'] / tot2:.0f}% of the step's kernel time by CUDA events; in the ncu "
                "window its per-launch mean against one launch of every steady-state training kernel gives the same share.",
                f"* roofline: {d['roofline']['achieved']:.0f} GB/s algorithmic = {100 * d['roofline']['frac']:.1f}% of the measured HBM copy peak "
                f"({d['roofline']['peak']} GB/s); gather+interaction {100 * d['roofline']['gather_interaction']['frac']:.1f}%."]
    open(os.path.join(P, "launches_r2.md"), "w").write("\n".join(out) + "\n")


# ------------------------------------------------------------------ full captures
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'launch__grid_size', 'launch__block_size', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    return rows[0], rows[1], rows[2:]


def metric_table(hdr, units, r):
    out = ["| metric | value |", "|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {r[i]} {units[i]} |")
    try:
        s_ = float(r[hdr.index('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum')].replace(',', ''))
        q_ = float(r[hdr.index('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum')].replace(',', ''))
        out.append(f"| sectors per global-load request | {s_ / q_:.2f} |")
    except Exception:
        pass
    return out


def gb(hdr, units, r, col):
    i = hdr.index(col)
    return float(r[i].replace(',', '')) * UNIT_GB[units[i]]


def full():
    out = ["# ncu --set full, round 2: the heavy kernels of the training step (B200, sm_100a)", "",
           "Commands (tools/profile_r2.sh, under gpurun, 1 GPU): `ncu --set full --clock-control none --import-source on -k regex:<kernel> "
           "-s 12 -c 2..3 -o gpurun_out/prof_<kernel>_r2 python bench.py --steps 4 --warmup 3 --working-set 4 --no-e2e --no-cpu-baseline "
           "--no-sweep --no-overlap-auc` (reports: gpurun_out/prof_*_r2.ncu-rep, not tracked).  Workload: B=65536 x 100 nnz raw ids, V_dim=64, "
           "every key owns a V row.", ""]
    traffic = {"source": "profiles/ncu_full_r2.md (ncu --set full, B=65536 x 100 nnz, V_dim=64; per launch)"}
    for rep, key in (("prof_bwd_r2.ncu-rep", "k_bwd_update<64>"), ("prof_fm_r2.ncu-rep", "k_fm_fast<64,emit>"),
                     ("prof_lookup_r2.ncu-rep", "k_lookup")):
        if not os.path.exists(os.path.join(G, rep)):
            continue
        hdr, units, rows = raw_rows(rep)
        r = rows[-1]
        out += [f"## `{nice(r[hdr.index('Kernel Name')])}`", ""] + metric_table(hdr, units, r) + [""]
        traffic[key] = {"dram_read_bytes": gb(hdr, units, r, 'dram__bytes_read.sum') * 1e9,
                        "dram_write_bytes": gb(hdr, units, r, 'dram__bytes_write.sum') * 1e9}
    out += ["Reading.",
            "* `k_fm_fast<64,emit>`: DRAM read 1.83 GB per launch against 1.73 GB algorithmic (N(4k+8)+16B) = **1.05x** (round 1: 2.29 GB = 1.29x). "
            "The per-load L2 policies did it: the 52 MB `{w,vrow}` view is loaded `evict_last` and now stays in L2 "
            "(`lts__t_sector_hit_rate` 7.5 % -> 22 %), the V rows stream through `evict_first`.",
            "* `k_bwd_update<64>`: 4.79 GB read + 3.53 GB written = 8.31 GB against 8.73 GB algorithmic, of which 1.68 GB are `p*XV` rows served from L2 "
            "(hit rate 47 %): no wasted re-reads beyond the 64-byte fetch granule of the 32-byte table entries.",
            "* `k_lookup`: one `LDG.E.256` per probe; 0.87 GB read for 6.5 M keys = 134 B per key: the 32-byte entry arrives in a 64-byte fetch "
            "granule and ~1.4 probes are needed per key at load factor 0.4.  Long-scoreboard stall 61 cycles per issue: pure memory latency at "
            "the DRAM activation rate of random granules (26 G/s); 0.34 ms (round 1) -> 0.25 ms by CUDA events.",
            "* No tensor-pipe instruction executes in any of them (`sm__pipe_tensor_cycles_active` 0): the path is ~0.5 FLOP/B."]
    open(os.path.join(P, "ncu_full_r2.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(P, "ncu_traffic_r2.json"), "w"), indent=1)


# ------------------------------------------------------------------ TMA A/B
def tma_ab():
    out = ["# K1 A/B: bulk-copy (cp.async.bulk + mbarrier) staged gather vs register-staged LDG.128 (round 2)", "",
           "north_star names \"TMA/shared-memory staging of the CSR row-block\"; VERDICT r1 #7 asked for one measured A/B.  The gathered rows "
           "are not a tile (every nnz names another 4k-byte table row), so the variant uses the 1-D bulk copy: every lane issues "
           "`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` for its nnz's row into a per-warp two-stage shared-memory ring "
           "(`csrc/kernels_fm_tma.cu`, engine kwarg `k1_tma=1`); the arithmetic then reads the rows with `LDS.128`.  Same inputs, same results "
           "(`tests/test_gpu_parity.py::test_k1_bulk_copy_variant_equals_register_staged_kernel`).", "",
           "## Timing (CUDA events inside bench.py, validation launches, B=65536 x 100 nnz, V_dim=64, kernels alone on the stream)", ""]
    ab = open(os.path.join(G, "ab4.log")).read().strip().splitlines() if os.path.exists(os.path.join(G, "ab4.log")) else []
    out += ["```"] + ab + ["```", "",
            "**LDG.128 (k_fm_fast<64,predict>): 0.316 ms = 83 % of the measured HBM peak; bulk copy (k_fm_tma<64>): 0.498 ms = 53 %.**  "
            "The register-staged kernel wins by 58 % and stays the default.", "", "## ncu --set full, both kernels", ""]
    for rep, title in (("prof_fm_r2.ncu-rep", "k_fm_fast<64,2,0> (LDG.128, training flavour of the same loop)"),
                       ("prof_fm_tma_r2.ncu-rep", "k_fm_tma<64,0> (bulk copy)")):
        if not os.path.exists(os.path.join(G, rep)):
            continue
        hdr, units, rows = raw_rows(rep)
        out += [f"### {title}", ""] + metric_table(hdr, units, rows[-1]) + [""]
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "difacto_b200", "lib", "kernels_fm_tma.o")],
                          capture_output=True, text=True).stdout.splitlines()
    keep = [l.rstrip() for l in sass if re.search(r"UBLKCP|SYNCS\.|LDS\.128|FENCE|Function :", l)][:26]
    sass2 = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "difacto_b200", "lib", "kernels_fm.o")],
                           capture_output=True, text=True).stdout
    m = re.search(r"Function : (\S*k_fm_fastILi64ELi0ELb0\S*)(.*?)(?=Function :|\Z)", sass2, re.S)
    keep2 = []
    if m:
        keep2 = [l.rstrip() for l in m.group(2).splitlines() if re.search(r"LDG\.E\.(128|64)|SHFL", l)][:14]
    out += ["Reading: the bulk-copy kernel is occupancy-limited by its 64 KB of shared memory per 4-warp CTA (`sm__warps_active` 18 % against 49 %), "
            "`UBLKCP` is a uniform-datapath instruction, so a warp issues its 32 row copies one after another, and a 256-byte transfer is too small to "
            "amortise the issue + mbarrier round trip; it also reads 2.19 GB from DRAM (no per-load L2 policy on the bulk path) against 1.83 GB.  The "
            "register-staged kernel already keeps 8 x 16 B per lane in flight at 64 registers.", "",
            "## SASS excerpts (cuobjdump -sass)", "", "`k_fm_tma<64,false>` — bulk copy + mbarrier:", "```"] + keep + ["```", "",
            "`k_fm_fast<64,0,false>` — register-staged gather (128-bit loads with an L2 cache-hint descriptor, shuffles for the broadcast):",
            "```"] + keep2 + ["```"]
    open(os.path.join(P, "k1_tma_ab.md"), "w").write("\n".join(out) + "\n")


# ------------------------------------------------------------------ NVLink + scaling
def nvlink_and_scaling():
    def parse(fn):
        tx = rx = 0
        gpu = 0
        for l in open(os.path.join(G, fn)):
            g = re.match(r'GPU (\d+):', l)
            if g:
                gpu = int(g.group(1))
            m = re.search(r'Link (\d+): Data (Tx|Rx): (\d+) KiB', l)
            if m and gpu == 0:          # GPU 0's links only (the listing may cover every GPU of the box)
                if m.group(2) == 'Tx':
                    tx += int(m.group(3))
                else:
                    rx += int(m.group(3))
        return tx * 1024 / 1e9, rx * 1024 / 1e9
    out = ["# NVLink traffic of the sharded step, measured with the links' own counters (round 2)", "",
           "`nvidia-smi nvlink -gt d` (cumulative data KiB per link; GPU 0's 18 links are summed) immediately before and after `torchrun ... bench.py "
           "--gpus N --steps 20 --warmup W` (ncu cannot wrap a multi-rank job).  A run contains 16 table warm-up steps (the first 8 with the "
           "feature-count push), W + 20 value steps, 10 profiled steps, W + 20 end-to-end steps, the small parity check and (final builds) the two "
           "Criteo-shaped sweep points.", "",
           "| GPUs | Tx GB (all 18 links of GPU 0) | Rx GB | measured per step / model for the run | model per full-size step (bench `roofline.nvlink`) | round-1 row exchange (model) |",
           "|---|---:|---:|---:|---:|---:|"]
    def model_bytes(world, B, N, U0, k):
        fr = (world - 1) / world
        return fr * (U0 * 12 + N * 16) + (world - 1) * B * (k + 2) * 4 + (world - 1) * B * (k + 1) * 4
    for tag, fns in (("nvl", ("bench_g2b.json",)), ("nvl8", ("bench_g8_final.json", "bench_g8.json"))):
        if not os.path.exists(os.path.join(G, tag + "_before.txt")):
            continue
        a, b = parse(tag + "_before.txt"), parse(tag + "_after.txt")
        d = next((bench_line(fn) for fn in fns if bench_line(fn)), None)
        if d is None:
            continue
        cfg = d["config"]
        if d.get("sweep") is not None or fns[0].endswith("final.json"):
            # the final runs: 16 table warm-up steps, warmup + steps value steps, max(4, steps/2) profiled steps, warmup + steps
            # end-to-end steps, then 8 + 3 + 6 + 4 steps per sweep point; the slices travel valued (16 bytes per nnz)
            full = 16 + (d["warmup"] + d["steps"]) * 2 + max(4, d["steps"] // 2)
            per = model_bytes(d["n_gpus"], cfg["batch_per_gpu"], cfg["batch_per_gpu"] * cfg["nnz_per_row"], cfg["unique_keys_per_batch"], cfg["V_dim"])
            total = full * per
            note = f"{full} full-size steps"
            for name, sw in (d.get("sweep") or {}).items():
                if "error" in sw:
                    continue
                total += 21 * model_bytes(d["n_gpus"], cfg["batch_per_gpu"], cfg["batch_per_gpu"] * 39, sw["unique_keys_per_batch"], sw["V_dim"])
                note += f" + 21 {name} steps"
            out.append(f"| {d['n_gpus']} | {b[0] - a[0]:.2f} | {b[1] - a[1]:.2f} | model for the whole run ({note}): {total / 1e9:.2f} GB | "
                       f"{per / 1e6:.0f} MB | {2 * d['roofline']['nvlink']['rows_exchange_model_bytes'] / 1e6:.0f} MB |")
        else:
            out.append(f"| {d['n_gpus']} (earlier build: binary slices, 8 bytes per nnz) | {b[0] - a[0]:.2f} | {b[1] - a[1]:.2f} | "
                       f"{(b[0] - a[0]) / 72 * 1e3:.0f} MB (72 steps) | "
                       f"{d['roofline']['nvlink']['bytes_out_per_gpu_per_step'] / 1e6:.0f} MB | "
                       f"{2 * d['roofline']['nvlink']['rows_exchange_model_bytes'] / 1e6:.0f} MB |")
    out += ["", "The counters follow the byte model of the protocol (slices of the batch's structure, (k+2) floats per row and owner one way, "
            "(k+1) floats per row and owner back) with 1.2-1.3x on top: the sub-CSR fill stores short runs (about a dozen 4-byte entries per "
            "row and owner), which the links carry in 32-byte granules, and the model leaves out the row pointers (0.5 MB per worker and owner) and "
            "the step headers / flags.  At 8 GPUs about 0.5 GB per full-size step leave a GPU instead of the 6.1 GB the row exchange of round 1 "
            "needs, i.e. ~120 GB/s of the 770 GB/s a GPU can send - NVLink is no longer what bounds the step."]
    open(os.path.join(P, "nvlink_r2.md"), "w").write("\n".join(out) + "\n")

    rows = []
    for fn in ("bench_full_final.json", "bench_full1.json"):
        if os.path.exists(os.path.join(G, fn)):
            rows.append(bench_line(fn))
            break
    for fn in ("bench_g2_final.json", "bench_g2b.json"):
        if os.path.exists(os.path.join(G, fn)):
            rows.append(bench_line(fn))
            break
    for fn in ("bench_g4_final.json", "bench_g4.json"):
        if os.path.exists(os.path.join(G, fn)) and bench_line(fn):
            rows.append(bench_line(fn))
            break
    for fn in ("bench_g8_final.json", "bench_g8.json"):
        if os.path.exists(os.path.join(G, fn)):
            rows.append(bench_line(fn))
            break
    v1 = rows[0]["value"]
    out = ["# Multi-GPU measurements, round 2 (builder runs; the driver's SCALE_r02.json is the judged record)", "",
           "`python bench.py` / `torchrun --nproc-per-node N bench.py --gpus N --steps 20 --warmup 3`; B=65536 x 100 nnz raw ids per GPU, V_dim=64, "
           "every key owns a V row; the NVLink-sharded store of csrc/shard.cu (no NCCL call inside a step).", "",
           "| GPUs | value M ex/s | ms/step | e2e M ex/s | efficiency v_N/(N v_1) | phases ms (rank 0) | phases sum / step | parity |",
           "|---|---:|---:|---:|---:|---|---:|---|"]
    for d in rows:
        ph = d.get("phases_ms_per_step_rank0")
        out.append(f"| {d['n_gpus']} | {d['value'] / 1e6:.1f} | {d['ms_per_step']:.3f} | {d['e2e']['value'] / 1e6:.1f} | "
                   f"{d['value'] / (d['n_gpus'] * v1):.2f} | "
                   + (", ".join(f"{k.replace('shard_', '')} {v:.2f}" for k, v in ph.items()) if ph else "-") + " | "
                   + (f"{d['phases_sum_ms'] / d['ms_per_step']:.2f}" if ph else "-") + " | "
                   + (("ok, max |dw| %.1e" % d["parity"]["max_abs_err"]["w"]) if d.get("parity") and d["parity"].get("ok") else "-") + " |")
    out += ["", "Round 1 (row exchange): 25.5 / 20.1 / 33.7 / 61.1 M ex/s at 1 / 2 / 4 / 8 GPUs (efficiency 0.39 / 0.33 / 0.30).  The phases are timed "
            "with CUDA events on the stream each runs on; their sum exceeds the step time because worker-side phases of step t+1 overlap the owner-side "
            "update of step t (and the owner waits inside `owner_updates` for the workers' p*XV)."]
    out += ["", "## Criteo-shaped sweep points on the same ranks (`sweep` object of the N-GPU line; 6 timed steps each)", "",
            "| GPUs | config | value M ex/s | ms/step | phases ms (rank 0) |", "|---|---|---:|---:|---|"]
    for d in rows:
        for name, sw in (d.get("sweep") or {}).items():
            if not name.startswith("criteo39") or "error" in sw or "allV" in name:
                continue
            ph = sw.get("phases_ms_per_step_rank0") or sw.get("stages_ms_per_step") or {}
            out.append(f"| {d['n_gpus']} | {name} | {sw['value'] / 1e6:.1f} | {sw['ms_per_step']:.3f} | "
                       + ", ".join(f"{k.replace('shard_', '')} {v:.2f}" for k, v in ph.items()) + " |")
    out += ["", "The Criteo-shaped steps are 4x smaller (39 nnz per row) and sub-millisecond on one GPU; across GPUs the four dependent hand-offs "
            "of a step and its ~80 launches per rank no longer hide behind the bandwidth-bound kernels, so these shapes scale worse than the "
            "100-nnz headline shape (0.24-0.30 efficiency at 8 GPUs)."]
    open(os.path.join(P, "scaling_r2.md"), "w").write("\n".join(out) + "\n")
    json.dump(rows[0], open(os.path.join(P, "bench_r2_1gpu.json"), "w"))
    for d in rows[1:]:
        json.dump(d, open(os.path.join(P, f"bench_r2_{d['n_gpus']}gpu.json"), "w"))


def shard_launches():
    """per-kernel device time of ONE rank-step of the sharded store (8 ranks as 8 engines on one GPU, tools/shard_local_profile.py)"""
    fn = os.path.join(G, "shard_launches8.csv")
    if not os.path.exists(fn):
        return
    rows = []
    for x in csv.DictReader([l for l in open(fn) if not l.startswith('==')]):
        try:
            rows.append((x['Kernel Name'], float(x['Metric Value'].replace(',', '')) / 1e3))
        except Exception:
            pass
    S = 8
    idx = [i for i, (k, _) in enumerate(rows) if 'k_rev_keys' in k]
    starts = [idx[i] for i in range(0, len(idx), S)]
    last = rows[starts[-2]:]          # the two timed steps (the four before them warm the table up)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in last:
        a = agg[nice(k)]
        a[0] += 1
        a[1] += v
    n = 2 * S
    tot = sum(v[1] for v in agg.values())
    out = ["# Launch list of the NVLink-sharded step, per rank and step (round 2)", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none python tools/shard_local_profile.py --ranks 8 --steps 2`: ncu must not wrap "
           "a multi-rank job, so the 8 ranks are 8 engines of ONE process on ONE GPU (same kernels, same mailbox traffic, device-local instead of "
           "NVLink; one host thread interleaves the enqueue phases, so every poller's flag is already set when ncu serialises the kernels).  "
           "B=65536 x 100 nnz per rank, V_dim=64; averages over the 16 rank-steps of the two timed collective steps; cold-cache, serialised times "
           "(compare shares).", "",
           f"Sum: {tot / n / 1e3:.2f} ms of kernel time per rank-step (the fused single-GPU step: 3.06 ms in profiles/launches_r2.md; run "
           "concurrently on the streams of a step the 8 local ranks take 3.1 ms per rank-step on this one GPU, `--time`).", "",
           "| kernel | launches / rank-step | us / rank-step | share |", "|---|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
        out.append(f"| `{k}` | {c / n:.2f} | {t / n:.1f} | {100 * t / tot:.1f} % |")
    out += ["", "Owner side: `k_bwd_update<64,shard>` x 8 (one Update per worker, rank order) costs what the single-GPU update costs (1.57 ms); "
            "`k_shard_lookup` (slots + sharing stamps, on the lookup stream beside the previous step's update) and `k_shard_pull` replace the fused "
            "step's `k_lookup`; `k_fm_fast<64,partial>` is K1 plus the (k+2)-float partial rows.  Worker side: the localizer's sort, the slice "
            "scatter / sub-CSR kernels (`k_shard_scatter`, `k_shard_rowcount`, `k_shard_rowscan`, `k_shard_fill`) and `k_shard_reduce`."]
    open(os.path.join(P, "shard_launches_r2.md"), "w").write("\n".join(out) + "\n")


def test_logs():
    for src, dst in (("t_full_r2.log", "pytest_gpu_1gpu_r2.log"), ("t_2gpu_r2.log", "pytest_gpu_2gpu_r2.log")):
        if os.path.exists(os.path.join(G, src)):
            open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())


if __name__ == "__main__":
    shard_launches()
    test_logs()
    launches()
    full()
    tma_ab()
    nvlink_and_scaling()
    print("profiles written:", sorted(f for f in os.listdir(P) if "r2" in f or "tma" in f))
