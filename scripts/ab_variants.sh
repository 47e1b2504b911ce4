#!/bin/bash
# A/B of the kernel variants built by scripts/build_variants.py on the GPU box (default workload, device-resident inputs).
# usage: scripts/ab_variants.sh [tags...]   -> gpurun_out/ab_variants.txt
mkdir -p gpurun_out
tags="$@"
[ -z "$tags" ] && tags=$(ls build/variants/libtt_b200_*.so | sed 's/.*libtt_b200_\(.*\)\.so/\1/')
for rep in ${TT_AB_REPS:-1 2}; do
  for t in $tags; do
    TT_B200_LIB=$PWD/build/variants/libtt_b200_$t.so python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $TT_BENCH_EXTRA 2>gpurun_out/ab_$t.err | tail -1 \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', 'rep$rep', 'value %.4g' % d['value'], 'kernel_ms %.2f' % d['roofline']['kernel_ms'], 'hist', d['histogram_sum'])" \
      | tee -a gpurun_out/ab_variants.txt
  done
done
