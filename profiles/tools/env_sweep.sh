#!/bin/bash
# usage: env_sweep.sh VAR v1 v2 ...  -- runs the default bench with VAR set to each value; prints kernel times + checksum
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernel_ms']
print('$var=$v', 'prove %.1f us  verify %.1f us  value %.3f G/s  e2e %.3f G/s' % (k['prove_kernel'] * 1e3, k['verify_kernel'] * 1e3, d['value'] / 1e9, d['e2e']['value'] / 1e9),
      'checksum', d['outcome']['proof_byte_checksum'], 'accept', d['outcome']['verified_accept'])
"
done
