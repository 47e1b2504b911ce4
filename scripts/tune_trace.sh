#!/bin/bash
# A/B on the GPU box: current working tree vs the last commit, default workload
python -m turbulence_tracing_b200.build --force > /dev/null
for rep in 1 2; do
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('working tree', d['value'], d['roofline']['kernel_ms'])"
done
python -m pytest tests -m gpu -q -k "variants_agree or trace_fp32 or grf129 or full_size" 2>&1 | tail -2
