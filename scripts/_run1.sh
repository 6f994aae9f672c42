mkdir -p gpurun_out
python scripts/perf_r2.py >> gpurun_out/perf_r2a.jsonl 2>> gpurun_out/perf_r2a.err
for lib in build_variants/libgx_nosph.so build_variants/libgx_sphmw.so; do
  GALAX_B200_LIB=$lib python scripts/perf_r2.py >> gpurun_out/perf_r2a.jsonl 2>> gpurun_out/perf_r2a.err
done
cat gpurun_out/perf_r2a.jsonl; tail -3 gpurun_out/perf_r2a.err
python -m pytest tests -m gpu -x -q -k "not c1_strict and not c2_dopri8 and not full_size" 2>&1 | tail -15
