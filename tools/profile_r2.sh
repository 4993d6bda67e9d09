#!/bin/bash
# tools/profile_r2.sh -- the ncu passes of B200_PROFILING.md for round 2 (run under gpurun, 1 GPU); raw outputs land in
# gpurun_out/, tools/make_profiles.py condenses them into profiles/
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 4 --warmup 3 --working-set 4 --no-e2e --no-cpu-baseline --no-sweep --no-overlap-auc"
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $BENCH > gpurun_out/ncu_launch_bench.log 2>&1
# the dominant kernels, full sets
ncu --set full --clock-control none --import-source on -k regex:k_bwd_update -s 12 -c 2 -o gpurun_out/prof_bwd_r2 -f $BENCH > gpurun_out/ncu_full_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fm_fast -s 12 -c 3 -o gpurun_out/prof_fm_r2 -f $BENCH > gpurun_out/ncu_full_fm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lookup -s 12 -c 2 -o gpurun_out/prof_lookup_r2 -f $BENCH > gpurun_out/ncu_full_lookup.log 2>&1
# the bulk-copy A/B of K1 (validation launches only)
ncu --set full --clock-control none --import-source on -k regex:k_fm_tma -s 2 -c 2 -o gpurun_out/prof_fm_tma_r2 -f $BENCH --engine-kw k1_tma=1 > gpurun_out/ncu_full_fm_tma.log 2>&1
ls -la gpurun_out/*.ncu-rep
