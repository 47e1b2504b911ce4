#!/bin/bash
# round 2, call 15: compute-sanitizer over the extended target (privatised detector kernel incl. counter overflow, aux grid, staged upload)
mkdir -p gpurun_out
python scripts/sanitizer_target.py > gpurun_out/sanitizer_plain.log 2>&1; tail -1 gpurun_out/sanitizer_plain.log
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_target.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_$tool.log | tail -1) / $(tail -1 gpurun_out/r02_sanitizer_$tool.log | cut -c1-80)"
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1
echo "smoke memcheck: $(grep 'ERROR SUMMARY' gpurun_out/r02_sanitizer_memcheck_smoke.log | tail -1)"
