#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one `--set full` capture per hot kernel.
# usage (on the GPU box): bash scripts/ncu_capture.sh <tag>      -> gpurun_out/{launches,prof_*}_<tag>.*
tag=${1:-r2}
mkdir -p gpurun_out
# (the bench's own ncu sub-invocation and the other-config legs are switched off: this pass lists the launches of the
#  headline loop -- device-resident steps, end-to-end steps, peak probes, the HBM leg)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-ncu > gpurun_out/bench_under_ncu_${tag}.log 2>&1
for k in fixed fixed_bovy dopri8 dopri8_1000; do
  kn=k_integrate_fixed; N=1212416
  case $k in dopri8*) kn=k_integrate_dopri8; N=303104;; esac
  N=$N ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -f -o gpurun_out/prof_${k}_${tag} \
      python scripts/ncu_one.py $k > gpurun_out/ncu_${k}_${tag}.log 2>&1
done
ls -la gpurun_out/*_${tag}*
