#!/bin/bash
# usage: poly_sweep.sh variants/lib_*.so -- runs bench.py --workload poly with each library build, prints the kernel times
for lib in "$@"; do
  PB_LIB=$lib python bench.py --workload poly --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_ms']
print('$lib', ' '.join('%s %.1f us' % (n.split()[0], v * 1e3) for n, v in k.items()), ' fused frac %.3f' % d['fused_roofline']['frac'])
"
done
