#!/bin/bash
# round 2, call 17: face-coefficient kernel with the passive quantities (configs[3]): GPU tests, occupancy A/B, bench c4
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -k "aux or c4 or kitchensink" 2>&1 | tail -4
out=gpurun_out/r02_ab_face_aux.txt; : > $out
for t in fa_new fa_mb2 fa_b64; do
  TT_B200_LIB=$PWD/build/variants/libtt_b200_$t.so python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', 'c4 value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'kernel_ms %.2f' % d['phases_ms']['trace_kernel'], d['checks']['ok'], d['config']['trace_path'])" | tee -a $out
done
python bench.py --workload c4 --no-face-grid --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('corner-grid', 'c4 value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'kernel_ms %.2f' % d['phases_ms']['trace_kernel'], d['checks']['ok'])" | tee -a $out
