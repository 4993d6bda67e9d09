#!/bin/bash
# tools/ab_lib.sh <lib.so|default> ... -- K1 / update stage times per V_dim for tuning builds (difacto_b200/build.py build_variant)
for lib in "$@"; do
  for k in 8 16 32 64 128; do
    if [ "$lib" = "default" ]; then unset DFB_LIB; else export DFB_LIB="$PWD/difacto_b200/lib/variants/$lib.so"; fi
    python bench.py --steps 6 --warmup 3 --working-set 4 --vdim $k --no-e2e --no-cpu-baseline --no-sweep 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$lib', 'k=$k', 'value=%.2fM' % (d['value']/1e6), 'K1pred=%.4f (%.2f)' % (d['roofline']['gather_interaction']['kernel_ms'], d['roofline']['gather_interaction']['frac']),
              'K1train=%.4f' % d['stages_ms_per_step']['fm'], 'upd=%.4f (%.2f)' % (d['stages_ms_per_step']['update'], d['roofline']['frac']), 'lookup=%.4f' % d['stages_ms_per_step']['lookup'], 'loc=%.4f' % d['stages_ms_per_step']['localize'])
"
  done
done
