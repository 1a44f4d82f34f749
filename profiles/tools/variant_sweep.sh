#!/bin/bash
# usage: variant_sweep.sh variants/lib_*.so  -- runs the default bench with each library build (PB_LIB override) and prints
# the prover / verifier kernel times and the outcome checksum (which must not change between variants)
for lib in "$@"; do
  PB_LIB=$lib python bench.py --steps 30 --warmup 5 --quick 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['roofline']['kernel_ms']
print('$lib', 'prove %.1f us  verify %.1f us  value %.3f G/s' % (k['prove_kernel'] * 1e3, k['verify_kernel'] * 1e3, d['value'] / 1e9),
      'checksum', d['outcome']['proof_byte_checksum'], 'accept', d['outcome']['verified_accept'])
"
done
