#!/bin/bash
# tuning sweep run on the GPU box: rebuild trace.cu with different settings and time the default workload
for pf in 0 2 4 8; do
  TT_NVCC_EXTRA="-DTT_EVENT_PREFETCH=$pf" python -m turbulence_tracing_b200.build --force > /dev/null
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prefetch=$pf', d['value'], d['roofline']['kernel_ms'])"
done
for mb in 4 6; do
  TT_NVCC_EXTRA="-DTT_EVENT_MIN_BLOCKS=$mb" python -m turbulence_tracing_b200.build --force > /dev/null
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minblocks=$mb', d['value'], d['roofline']['kernel_ms'])"
done
python -m turbulence_tracing_b200.build --force > /dev/null
