#!/bin/bash
# A/B on the GPU box: in-tree build first, then forced rebuilds with experiment flags (default workload, device-resident
# inputs).  Prepared, not yet measured (NOTES.md): TT_EVENT_LEAN, TT_AXES_RCP, TT_EVENT_BLOCK.
run() {
  for rep in 1 2; do
    python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['kernel_ms'])"
  done
}
run "in-tree build"
for flags in "-DTT_EVENT_LEAN=1" "-DTT_EVENT_BLOCK=64 -DTT_EVENT_MIN_BLOCKS=10" "-DTT_EVENT_BLOCK=256 -DTT_EVENT_MIN_BLOCKS=2"; do
  TT_NVCC_EXTRA="$flags" python -m turbulence_tracing_b200.build --force > /dev/null 2>&1 && run "$flags"
done
TT_NVCC_EXTRA="-DTT_AXES_RCP=1" python -m turbulence_tracing_b200.build --force > /dev/null 2>&1 && python scripts/bench_rectilinear.py | head -2
python -m turbulence_tracing_b200.build --force > /dev/null
python -m pytest tests -m gpu -q -k "variants_agree or trace_fp32 or grf129 or full_size or rectilinear" 2>&1 | tail -2
