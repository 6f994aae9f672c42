bash scripts/ncu_capture.sh r2k > gpurun_out/ncu_capture_r2k.log 2>&1
for k in fixed fixed_bovy dopri8 dopri8_1000; do
  u=""; case $k in fixed*) u=1.212416e10;; esac
  python scripts/ncu_summary.py gpurun_out/prof_${k}_r2k.ncu-rep $u > gpurun_out/ncu_summary_${k}_r2k.txt 2>&1
  ncu -i gpurun_out/prof_${k}_r2k.ncu-rep --page source --csv --print-source sass > gpurun_out/sass_${k}_r2k.csv 2>/dev/null
done
rm -f gpurun_out/prof_fixed_r2k.ncu-rep gpurun_out/prof_fixed_bovy_r2k.ncu-rep gpurun_out/prof_dopri8_r2k.ncu-rep
gzip -f gpurun_out/sass_*_r2k.csv
ls -la gpurun_out | head -30
