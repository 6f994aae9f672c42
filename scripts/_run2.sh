for v in mix34b mix58 mix1116 mix1316; do
GALAX_B200_LIB=build_variants/libgx_$v.so python scripts/perf_r2.py k2 2>&1 | tail -1 | cut -c1-80
done
