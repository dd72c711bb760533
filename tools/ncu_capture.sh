#!/bin/bash
# ncu --set full captures of the dominant kernels, skipping bench.py's first launches (the stats pass and the
# parity spot check).  Run through gpurun; reports land in gpurun_out/.
R=${1:-r02}
mkdir -p gpurun_out
cap() { # name, kernel regex, bench args...
  local name=$1 rx=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip 2 --launch-count 1 -f \
    -o gpurun_out/prof_${R}_${name} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-multi-query --no-replay "$@" > gpurun_out/${R}_ncu_${name}.log 2>&1
  # gpurun_out/ is capped at 64 MiB and a report with sources is ~20 MB: keep the text exports, and the report
  # itself only when asked (KEEP_REP=1)
  ncu -i gpurun_out/prof_${R}_${name}.ncu-rep --page raw --csv > gpurun_out/prof_${R}_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${R}_${name}.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${R}_${name}_sass.csv 2>/dev/null
  ncu -i gpurun_out/prof_${R}_${name}.ncu-rep --page details > gpurun_out/prof_${R}_${name}_details.txt 2>/dev/null
  if [ "${KEEP_REP:-0}" != "1" ]; then rm -f gpurun_out/prof_${R}_${name}.ncu-rep; fi
}
KEEP_REP=1 cap fxn_headline expand_fxn_kernel
cap fxn_cfg2 expand_fxn_kernel --workload cfg2
cap fxn_cfg3 expand_fxn_kernel --workload cfg3
cap deal_cfg4 expand_deal_kernel --workload cfg4
ls -la gpurun_out/ | tail -20
