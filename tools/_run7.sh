mkdir -p gpurun_out
( timeout 300 python tools/dmma_check.py ) > gpurun_out/r7_dmma.log 2>&1
cat gpurun_out/r7_dmma.log | tail -4
( timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q ) > gpurun_out/r7_tests.log 2>&1
tail -3 gpurun_out/r7_tests.log
