mkdir -p gpurun_out
( time timeout 300 python tools/dmma_check.py ) > gpurun_out/r4_dmma.log 2>&1
cat gpurun_out/r4_dmma.log | tail -6
( time timeout 300 python tools/walker_trace.py 20000 1184 1 ) > gpurun_out/r4_trace.log 2>&1
tail -12 gpurun_out/r4_trace.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r4_tests.log 2>&1
tail -4 gpurun_out/r4_tests.log
( time timeout 900 python bench.py --steps 2 --warmup 1 ) > gpurun_out/r4_bench.log 2>&1
tail -c 600 gpurun_out/r4_bench.log
