#!/bin/bash
# DRAM traffic of the save path: K3 with 10 and 1000 saves, K2 with 101 saves, [N,T,3] layout.
# usage (GPU box): bash scripts/ncu_traffic.sh <tag>  ->  gpurun_out/traffic_<tag>.csv
tag=${1:-r2}
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
: > gpurun_out/traffic_${tag}.csv
for spec in "dopri8 303104 k_integrate_dopri8" "dopri8_1000 303104 k_integrate_dopri8" "fixed_101 1212416 k_integrate_fixed"; do
  set -- $spec
  echo "# $1 N=$2" >> gpurun_out/traffic_${tag}.csv
  N=$2 ncu --metrics $M --clock-control none -k regex:$3 -c 1 --csv python scripts/ncu_one.py $1 2>/dev/null | grep -E "dram__|gpu__time" >> gpurun_out/traffic_${tag}.csv
done
cat gpurun_out/traffic_${tag}.csv
