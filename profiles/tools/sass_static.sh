#!/bin/bash
# usage: sass_static.sh <mangled-name-fragment> [library]  -- static SASS opcode mix of one kernel in the built library.
# The prover and verifier are straight-line, so their static mix is within a few percent of the executed one: a change can
# be sized here (no GPU needed) before it is measured.  Example: sass_static.sh prove_kernelINS_16ProverWideTablesELb0
frag=$1; lib=${2:-plonk.c_b200/libplonk_b200.so}
cuobjdump -sass "$lib" 2>/dev/null | awk -v f="$frag" '/Function : /{on = index($0, f) > 0} on' \
  | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9]+\s+)?//' | awk '{print $1}' \
  | sed -E 's/^(IMAD\.(HI|WIDE|MOV|IADD|SHL|X)).*/\1/; t; s/\..*//' | sort | uniq -c | sort -rn | awk '{t += $1; print} END {print t, "total"}'
