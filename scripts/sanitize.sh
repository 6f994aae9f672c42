#!/bin/bash
# compute-sanitizer over the GPU tests that reach every kernel (small sizes).  usage: bash scripts/sanitize.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
SEL='not c1_strict and not c2_dopri8 and not full_size and not c1_parity and not host_threads and not joint_mode_properties and not other_kernels_and_layouts'
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/memcheck_${tag}.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_${tag}.log
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_fixtures.py tests/test_gpu_abi_contract.py tests/test_gpu_epilogue.py tests/test_gpu_joint.py -m gpu -q -k "not host_threads and not joint_mode_properties and not other_kernels_and_layouts" > gpurun_out/racecheck_${tag}.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck_${tag}.log | tail -3
compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_fixtures.py tests/test_gpu_experimental.py tests/test_gpu_joint.py -m gpu -q -k "not joint_mode_properties" > gpurun_out/synccheck_${tag}.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/synccheck_${tag}.log | tail -3
