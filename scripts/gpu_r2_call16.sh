#!/bin/bash
# round 2, call 16: the two experiments prepared in round 1 and never measured: TT_EVENT_LEAN (corner-grid packed kernel: c3 through
# tt_trace, c4 with the passive quantities) and TT_AXES_RCP (rectilinear event kernel)
mkdir -p gpurun_out
out=gpurun_out/r02_ab_lean_axesrcp.txt; : > $out
for t in a_new e_lean; do
  for wl in "c3 --no-face-grid" "c4"; do
    TT_B200_LIB=$PWD/build/variants/libtt_b200_$t.so python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', '$wl', 'value %.4g' % d['value'], 'kernel_ms %.2f' % d['phases_ms']['trace_kernel'], d['checks']['ok'])" | tee -a $out
  done
done
for t in x_new x_rcp; do
  echo "== $t" | tee -a $out
  TT_B200_LIB=$PWD/build/variants/libtt_b200_$t.so python scripts/bench_rectilinear.py 2>&1 | grep "variant 0" | tee -a $out
done
