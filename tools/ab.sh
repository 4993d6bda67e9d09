#!/bin/bash
# tools/ab.sh -- quick A/B of engine knobs on the 1-GPU bench (stage timings only): prints one line per variant
for kw in "$@"; do
  python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --engine-kw "$kw" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$kw', 'value=%.2fM' % (d['value']/1e6), {k: round(v, 4) for k, v in d['stages_ms_per_step'].items()}, 'K1 predict ms', round(d['roofline']['gather_interaction']['kernel_ms'], 4))
"
done
