#!/bin/bash
# tools/ab32.sh -- A/B of the V_dim=32 forward kernel: product build (4 CTAs/SM, 64 registers, ~130 bytes of spills) against
# a tuning build without spills (3 CTAs/SM, 80 registers).  Build the variant first (here, no GPU needed):
#   python -c "from difacto_b200 import build; build.build_variant('fm32_mb3', ['-DDFB_FM_MINBLOCKS32=3'])"
# then run under gpurun: bash tools/ab32.sh.  Result (round 2): 0.2456 ms (product) vs 0.2548 ms (no spills) -> DESIGN.md section 3.
for lib in default fm32_mb3; do
  if [ "$lib" = "default" ]; then unset DFB_LIB; else export DFB_LIB="$PWD/difacto_b200/lib/variants/$lib.so"; fi
  python bench.py --steps 6 --warmup 3 --working-set 4 --vdim 32 --no-e2e --no-cpu-baseline --no-sweep 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$lib', 'value=%.2fM' % (d['value']/1e6), 'K1pred=%.4f (%.2f)' % (d['roofline']['gather_interaction']['kernel_ms'], d['roofline']['gather_interaction']['frac']), 'K1train=%.4f' % d['stages_ms_per_step']['fm'], 'upd=%.4f' % d['stages_ms_per_step']['update'])
"
done
