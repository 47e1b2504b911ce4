#!/bin/bash
# A/B of TT_EVENT_REUSE on the GPU box: in-tree build (=1) first, then a forced rebuild with =0
mkdir -p gpurun_out
run() {  # label
  for rep in 1 2; do
    timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 c3 f32', d['value'], d['roofline']['kernel_ms'])"
  done
  timeout 300 python bench.py --workload c2 --dtype float64 --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 c2 f64', d['value'], d['roofline']['kernel_ms'])"
  timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 c4 aux', d['value'], d['roofline']['kernel_ms'])"
}
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu8.log 2>&1; grep -n "passed\|failed" gpurun_out/pytest_gpu8.log
run reuse1 | tee gpurun_out/ab_reuse.log
TT_NVCC_EXTRA=-DTT_EVENT_REUSE=0 python -m turbulence_tracing_b200.build --force > /dev/null 2>&1
run reuse0 | tee -a gpurun_out/ab_reuse.log
