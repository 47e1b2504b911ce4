#!/bin/bash
# tuning sweep run on the GPU box: rebuild with different settings and time the default workload
for rep in 1 2; do
for opt in "" "-DTT_RCP_NEWTON=0"; do
  TT_NVCC_EXTRA="$opt" python -m turbulence_tracing_b200.build --force > /dev/null
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('opt=[$opt]', d['value'], d['roofline']['kernel_ms'])"
done
done
TT_NVCC_EXTRA="-DTT_RCP_NEWTON=0" python -m turbulence_tracing_b200.build --force > /dev/null
python -m pytest tests -m gpu -q -k "trace_fp32 or grf129 or c1_end or full_size or large_bundle" 2>&1 | tail -2
python -m turbulence_tracing_b200.build --force > /dev/null
