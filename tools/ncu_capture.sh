#!/bin/bash
# ncu --set full captures of the dominant kernels, skipping bench.py's first launch (the one with the
# stats counters enabled, which adds shared-memory atomics).  Run through gpurun; reports land in gpurun_out/.
R=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_reg_kernel --launch-skip 2 --launch-count 1 -f \
  -o gpurun_out/prof_${R}_reg python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_reg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_deal_kernel --launch-skip 2 --launch-count 1 -f \
  -o gpurun_out/prof_${R}_deal_cfg3 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_deal.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_reg_kernel --launch-skip 2 --launch-count 1 -f \
  -o gpurun_out/prof_${R}_reg_cfg3 python bench.py --workload cfg3 --kernel 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_reg_cfg3.log 2>&1
ls -la gpurun_out/*.ncu-rep
