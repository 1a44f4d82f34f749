#!/bin/bash
# Final ncu evidence of a round, digested ON the GPU box (the reports together exceed what gpurun brings back); usage:
#   gpurun -- 'bash profiles/tools/capture_all.sh v16'     -> gpurun_out/ncu_<name>_<tag>.{txt,json} (+ the cold headline report)
tag=$1
N21=2097152
cap() {   # name target regex extra-ncu-args items
  local name=$1 target=$2 regex=$3 extra=$4 items=$5
  python profiles/tools/ncu_target.py $target > gpurun_out/plain_$name.log 2>&1 || { echo "$name: plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:"$regex" $extra -o gpurun_out/prof_${name}_$tag python profiles/tools/ncu_target.py $target > gpurun_out/ncu_$name.log 2>&1
  python profiles/tools/ncu_digest.py gpurun_out/prof_${name}_$tag.ncu-rep --items $items --json gpurun_out/ncu_${name}_$tag.json > gpurun_out/ncu_${name}_$tag.txt 2>gpurun_out/digest_$name.err
  echo "$name: $(grep -c '^==' gpurun_out/ncu_${name}_$tag.txt) launches digested"
}
cap headline_cold headline "prove_kernel|verify_fast|verify_log" "-s 4 -c 2" $N21
cap headline_warm headline "prove_kernel|verify_fast|verify_log" "--cache-control none -s 4 -c 2" $N21
rm -f gpurun_out/prof_headline_warm_$tag.ncu-rep
cap headline_packed headline_packed "prove_kernel|verify_fast|verify_log|gather_done|done_offsets" "-s 8 -c 4" $N21
rm -f gpurun_out/prof_headline_packed_$tag.ncu-rep
cap headline_pair headline_pair "prove_kernel|verify_fast" "-s 4 -c 2" $N21
rm -f gpurun_out/prof_headline_pair_$tag.ncu-rep
# families: items per launch in launch order (ncu_target.py families): pairing 2^22, g1_mul 2^24, field x2 2^27, then 2^22 each
cap families families "pairing_kernel|g1_mul_kernel|field_op_kernel|field_pow_kernel|poly_binop_kernel|poly_divide_kernel|poly_mul_fast|poly_divide_fast|poly_divide_zh|poly_eval_fast|interpolate4|config2_kernel" "" 4194304,16777216,134217728,134217728,4194304
rm -f gpurun_out/prof_families_$tag.ncu-rep
ls -la gpurun_out | head -40
