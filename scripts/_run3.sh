bash scripts/ncu_capture.sh r2k > gpurun_out/ncu_capture_r2k.log 2>&1
bash scripts/ncu_traffic.sh r2k > gpurun_out/ncu_traffic_r2k.log 2>&1
tail -3 gpurun_out/ncu_capture_r2k.log; cat gpurun_out/traffic_r2k.csv | cut -c1-30,150-400
