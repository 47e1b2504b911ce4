#!/bin/bash
# tuning sweep run on the GPU box: rebuild trace.cu with different settings and time the default workload
for rep in 1 2; do
for opt in "" "-DTT_EXACT_LAMBDA"; do
  TT_NVCC_EXTRA="$opt" python -m turbulence_tracing_b200.build --force > /dev/null
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('opt=[$opt]', d['value'], d['roofline']['kernel_ms'])"
done
done
python -m turbulence_tracing_b200.build --force > /dev/null
