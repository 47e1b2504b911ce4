#!/bin/bash
# round 2, call 12: L2 prefetch of the face grid (planes ahead of the register prefetch) on configs[4] (1025^3) and configs[2]
mkdir -p gpurun_out
rm -f gpurun_out/ab_variants.txt
for wl in c5 c3; do
  echo "workload $wl" | tee -a gpurun_out/ab_variants.txt
  TT_BENCH_EXTRA="--no-extras --workload $wl" bash scripts/ab_variants.sh f_new f_pf2 f_pf4 f_pf8
done
cp gpurun_out/ab_variants.txt gpurun_out/r02_ab_face_prefetch.txt
