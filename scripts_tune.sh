#!/bin/bash
# tuning sweep run on the GPU box: rebuild trace.cu with different register targets and time c2
for mb in 3 4 5; do
  TT_NVCC_EXTRA="-DTT_TRACE_MIN_BLOCKS=$mb" python -m turbulence_tracing_b200.build --force > /dev/null
  python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minblocks=$mb', d['value'], d['roofline']['kernel_ms'])"
done
python -m turbulence_tracing_b200.build --force > /dev/null
