#!/bin/bash
# round 2, call 13: threads per CTA of the privatised detector kernel; CTAs per SM of the packed AUX kernel (configs[3])
mkdir -p gpurun_out
: > gpurun_out/r02_ab_hist_aux.txt
for t in h_new h_t384 h_t640 h_t768; do
  echo "== $t" | tee -a gpurun_out/r02_ab_hist_aux.txt
  TT_B200_LIB=$PWD/build/variants/libtt_b200_$t.so python scripts/diag_optics.py 2>&1 | grep "original order" | tee -a gpurun_out/r02_ab_hist_aux.txt
done
for t in a_new a_mb2 a_mb4; do
  for rep in 1 2; do
  TT_B200_LIB=$PWD/build/variants/libtt_b200_$t.so python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', 'c4 value %.4g' % d['value'], 'kernel_ms %.2f' % d['phases_ms']['trace_kernel'], d['checks']['ok'])" | tee -a gpurun_out/r02_ab_hist_aux.txt
  done
done
python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['phases_ms'], d['roofline'].get('detector_kernel'))"
